#!/bin/bash
# round-2 GPU batch B: full GPU test suite on the refactored build, default bench, prefetch / occupancy experiments
mkdir -p gpurun_out
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/r02b_pytest_gpu.log 2>&1
python __graft_entry__.py smoke > $O/r02b_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/r02b_bench_default.json 2> $O/r02b_bench_default.err
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
for pf in 0 1; do
  $B --workload efit_rect --prefetch $pf > $O/r02b_efit_rect_k2_pf$pf.json 2>> $O/r02b_err.log
  $B --workload efit_rect --ipusher 1 --prefetch $pf > $O/r02b_efit_rect_rk4_pf$pf.json 2>> $O/r02b_err.log
  $B --start spread --prefetch $pf > $O/r02b_vmec_spread_k2_pf$pf.json 2>> $O/r02b_err.log
  $B --workload west_soledge3x --prefetch $pf > $O/r02b_west_rk4_pf$pf.json 2>> $O/r02b_err.log
done
$B --prefetch 1 > $O/r02b_vmec_k2_pf1.json 2>> $O/r02b_err.log
$B --ipusher 1 --prefetch 1 > $O/r02b_vmec_rk4_pf1.json 2>> $O/r02b_err.log
for v in k2m5 k2m6; do
  export GORILLA_B200_LIB=$PWD/gorilla_b200/lib/libgorilla_b200_$v.so
  $B > $O/r02b_vmec_k2_$v.json 2>> $O/r02b_err.log
  for pf in 0 1; do $B --workload efit_rect --prefetch $pf > $O/r02b_efit_rect_k2_${v}_pf$pf.json 2>> $O/r02b_err.log; done
done
export GORILLA_B200_LIB=$PWD/gorilla_b200/lib/libgorilla_b200_rkm4.so
$B --ipusher 1 > $O/r02b_vmec_rk4_rkm4.json 2>> $O/r02b_err.log
for pf in 0 1; do $B --workload efit_rect --ipusher 1 --prefetch $pf > $O/r02b_efit_rect_rk4_rkm4_pf$pf.json 2>> $O/r02b_err.log; done
unset GORILLA_B200_LIB
tail -5 $O/r02b_pytest_gpu.log
