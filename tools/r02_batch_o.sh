#!/bin/bash
# round-2 batch O (8 GPUs): weak scaling of the headline workload and config 5 (strong, 1e7 alphas) through the library's NCCL communicator
mkdir -p gpurun_out
O=gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
$TR --steps 5 --warmup 3 --no-variants > $O/r02o_bench_8gpu.json 2> $O/r02o_bench_8gpu.err
$TR --steps 3 --warmup 2 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 > $O/r02o_bench_config5_strong_k2_8gpu.json 2>> $O/r02o_err.log
$TR --steps 2 --warmup 1 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 --ipusher 1 > $O/r02o_bench_config5_strong_rk4_8gpu.json 2>> $O/r02o_err.log
for f in $O/r02o_bench_*.json; do echo $f; cut -c1-200 $f; done; tail -3 $O/r02o_err.log; tail -3 $O/r02o_bench_8gpu.err
