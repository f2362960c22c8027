#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(time timeout 600 python -m pytest tests/test_diag_and_resort.py tests/test_gpu_parity.py tests/test_rk_pusher.py tests/test_strong_electric_field.py -m gpu -q) > $O/r02d_pytest_gpu.log 2>&1
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --prefetch 0"
for ga in 0 1; do
  $B --workload efit_rect --gather $ga > $O/r02d_efit_rect_k2_ga$ga.json 2>> $O/r02d_err.log
  $B --workload efit_rect --ipusher 1 --gather $ga > $O/r02d_efit_rect_rk4_ga$ga.json 2>> $O/r02d_err.log
  $B --gather $ga > $O/r02d_vmec_k2_ga$ga.json 2>> $O/r02d_err.log
  $B --ipusher 1 --gather $ga > $O/r02d_vmec_rk4_ga$ga.json 2>> $O/r02d_err.log
  $B --start spread --gather $ga > $O/r02d_vmec_spread_k2_ga$ga.json 2>> $O/r02d_err.log
  $B --workload west_soledge3x --gather $ga > $O/r02d_west_rk4_ga$ga.json 2>> $O/r02d_err.log
  $B --workload west_soledge3x --ipusher 2 --gather $ga > $O/r02d_west_k2_ga$ga.json 2>> $O/r02d_err.log
done
tail -5 $O/r02d_pytest_gpu.log
