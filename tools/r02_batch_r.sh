#!/bin/bash
# round-2 batch R (4 GPUs): the missing point of the 1/2/4/8 table; then synccheck of the bulk-copy kernels with a larger barrier table
mkdir -p gpurun_out
O=gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4"
$TR --steps 5 --warmup 3 --no-variants --no-cpu-baseline > $O/r02r_bench_4gpu.json 2> $O/r02r_bench_4gpu.err
$TR --steps 3 --warmup 2 --no-variants --no-cpu-baseline --scaling strong --total-particles 10000000 > $O/r02r_bench_config5_strong_k2_4gpu.json 2>> $O/r02r_err.log
(CUDA_VISIBLE_DEVICES=0 timeout 600 compute-sanitizer --error-exitcode 7 --launch-timeout 0 --tool synccheck --num-cuda-barriers 400000 python -m pytest -m gpu -q tests/test_diag_and_resort.py -k "gather_and_prefetch") > $O/r02r_synccheck_bulk.log 2>&1; echo "synccheck_bulk rc=$?" > $O/r02r_summary.log
for f in $O/r02r_bench_*.json; do echo $f; cut -c1-200 $f; done; cat $O/r02r_summary.log; grep -E "ERROR SUMMARY|passed|failed|Warning" $O/r02r_synccheck_bulk.log | tail -4
