#!/bin/bash
# round-2 measurement batch J (2 GPUs): the 2-rank test through the library communicator, weak and strong (config 5) scaling
mkdir -p gpurun_out
O=gpurun_out
(time timeout 600 python -m pytest tests/test_multi_gpu_gloo.py tests/test_diag_and_resort.py -m gpu -q) > $O/r02j_pytest_2gpu.log 2>&1
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
$TR --steps 5 --warmup 3 --no-variants > $O/r02j_bench_2gpu.json 2> $O/r02j_bench_2gpu.err
$TR --impl reference --steps 2 --warmup 1 > $O/r02j_bench_2gpu_reference.json 2>> $O/r02j_err.log
$TR --steps 3 --warmup 2 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 > $O/r02j_bench_config5_strong_k2_2gpu.json 2>> $O/r02j_err.log
$TR --steps 2 --warmup 1 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 --ipusher 1 > $O/r02j_bench_config5_strong_rk4_2gpu.json 2>> $O/r02j_err.log
$TR --steps 1 --warmup 1 --no-variants --no-cpu-baseline --no-e2e --scaling strong --total-particles 10000000 --poly-order 4 > $O/r02j_bench_config5_strong_k4_2gpu.json 2>> $O/r02j_err.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 > $O/r02j_bench_config5_strong_k2_1gpu.json 2>> $O/r02j_err.log
tail -3 $O/r02j_pytest_2gpu.log; for f in $O/r02j_bench_*.json; do echo $f; cut -c1-250 $f; done; tail -5 $O/r02j_err.log
