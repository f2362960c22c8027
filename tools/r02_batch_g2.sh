#!/bin/bash
# round-2 batch G2 (2 GPUs): the driver's multi-GPU launch of both bench arms on the final build + the 2-rank test through the library communicator
mkdir -p gpurun_out
O=gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2"
$TR --steps 5 --warmup 3 --no-variants > $O/r02g_bench_2gpu.json 2> $O/r02g_bench_2gpu.err
$TR --steps 2 --warmup 1 --impl reference > $O/r02g_bench_2gpu_reference.json 2> $O/r02g_bench_2gpu_reference.err
(timeout 600 python -m pytest -m gpu -q tests/test_multi_gpu_gloo.py) > $O/r02g_pytest_2gpu.log 2>&1
tail -n 2 $O/r02g_pytest_2gpu.log
for f in $O/r02g_bench_2gpu.json $O/r02g_bench_2gpu_reference.json; do cut -c1-230 $f; done
