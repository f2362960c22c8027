#!/bin/bash
# round-2 final batch 3 (2 GPUs): the 2-rank test through the library communicator and a short weak-scaling bench of the final build
mkdir -p gpurun_out
O=gpurun_out
(time timeout 60 python -m pytest tests/test_multi_gpu_gloo.py -m gpu -q) > $O/r02fin_pytest_2gpu.log 2>&1
(time timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-variants) > $O/r02fin_bench_2gpu.json 2> $O/r02fin_bench_2gpu.err
tail -n 4 $O/r02fin_pytest_2gpu.log | cut -c1-200; cut -c1-300 $O/r02fin_bench_2gpu.json; tail -n 3 $O/r02fin_bench_2gpu.err
