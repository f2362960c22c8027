#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_diag_and_resort.py tests/test_precomp_modes.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q) > $O/r02e_pytest_gpu.log 2>&1
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --prefetch 0"
for ga in 0 1; do
  $B --workload efit_rect --gather $ga > $O/r02e_efit_rect_k2_ga$ga.json 2>> $O/r02e_err.log
  $B --workload efit_rect --ipusher 1 --gather $ga > $O/r02e_efit_rect_rk4_ga$ga.json 2>> $O/r02e_err.log
  $B --workload west_soledge3x --gather $ga > $O/r02e_west_rk4_ga$ga.json 2>> $O/r02e_err.log
  $B --workload west_soledge3x --ipusher 2 --gather $ga > $O/r02e_west_k2_ga$ga.json 2>> $O/r02e_err.log
done
$B --gather 1 > $O/r02e_vmec_k2_ga1.json 2>> $O/r02e_err.log
$B --start spread --gather 1 > $O/r02e_vmec_spread_k2_ga1.json 2>> $O/r02e_err.log
$B --workload efit_flux --gather 0 > $O/r02e_efit_flux_k2_ga0.json 2>> $O/r02e_err.log
$B --workload efit_flux --gather 1 > $O/r02e_efit_flux_k2_ga1.json 2>> $O/r02e_err.log
$B --workload efit_flux --poly-order 2 --gather 0 --particles 300000 > /dev/null 2>&1
python bench.py --workload efit_rect --gather 1 --prefetch 0 --steps 3 --warmup 3 --no-cpu-baseline --no-variants > $O/r02e_efit_rect_k2_ga1_e2e.json 2>> $O/r02e_err.log
tail -5 $O/r02e_pytest_gpu.log
