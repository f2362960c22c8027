#!/bin/bash
# round-2 profiling pass A: ncu --set full captures (source-level) of the kernels VERDICT names
set -x
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
$NCU -o gpurun_out/r02a_efit_rect_k2 python bench.py --workload efit_rect --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02a_efit_rect_k2.log 2>&1
$NCU -o gpurun_out/r02a_vmec_k4 python bench.py --poly-order 4 --particles 300000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02a_vmec_k4.log 2>&1
$NCU -o gpurun_out/r02a_vmec_rk4 python bench.py --ipusher 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02a_vmec_rk4.log 2>&1
python bench.py --workload efit_rect --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_efit_rect_k2.json 2> gpurun_out/r02a_bench_efit_rect_k2.err
python bench.py --workload efit_rect --ipusher 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_efit_rect_rk4.json 2> gpurun_out/r02a_bench_efit_rect_rk4.err
python bench.py --poly-order 4 --particles 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_vmec_k4.json 2> gpurun_out/r02a_bench_vmec_k4.err
python bench.py --ipusher 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_vmec_rk4.json 2> gpurun_out/r02a_bench_vmec_rk4.err
ls -la gpurun_out
