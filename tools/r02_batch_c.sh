#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(time python -m pytest tests -m gpu -q) > $O/r02c_pytest_gpu.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --prefetch 0"
$B --workload efit_rect --particles 4000000 > $O/r02c_efit_rect_k2_4M.json 2>> $O/r02c_err.log
tail -5 $O/r02c_pytest_gpu.log
