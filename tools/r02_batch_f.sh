#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --prefetch 0"
export GORILLA_B200_LIB=$PWD/gorilla_b200/lib/libgorilla_b200_ldgsts.so
timeout 300 python -m pytest tests/test_diag_and_resort.py -m gpu -q -k gather > $O/r02f_pytest.log 2>&1
$B --workload efit_rect --gather 1 > $O/r02f_efit_rect_k2_ldgsts.json 2>> $O/r02f_err.log
$B --gather 1 > $O/r02f_vmec_k2_ldgsts.json 2>> $O/r02f_err.log
$B --workload west_soledge3x --ipusher 2 --gather 1 > $O/r02f_west_k2_ldgsts.json 2>> $O/r02f_err.log
$B --workload efit_rect --ipusher 1 --gather 1 > $O/r02f_efit_rect_rk4_ldgsts.json 2>> $O/r02f_err.log
tail -3 $O/r02f_pytest.log
