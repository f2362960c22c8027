#!/bin/bash
# round-2 batch U: cooperative gather with 15 of the 22 record pieces staged and FOUR CTAs per SM (variant library c15)
# against the full-record form at three CTAs per SM (main library)
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/gorilla_b200/lib/libgorilla_b200_c15.so
(timeout 600 python -m pytest -m gpu -q -x tests/test_diag_and_resort.py -k "gather_and_prefetch") > $O/r02u_pytest_main.log 2>&1
(GORILLA_B200_LIB=$V timeout 600 python -m pytest -m gpu -q -x tests/test_diag_and_resort.py -k "gather_and_prefetch") > $O/r02u_pytest_c15.log 2>&1
tail -2 $O/r02u_pytest_main.log $O/r02u_pytest_c15.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --gather 2"
for lib in main c15; do
  if [ $lib = c15 ]; then export GORILLA_B200_LIB=$V; else unset GORILLA_B200_LIB; fi
  $B > $O/r02u_vmec_k2_$lib.json 2>> $O/r02u_err.log
  $B --workload efit_rect > $O/r02u_efit_rect_k2_$lib.json 2>> $O/r02u_err.log
  $B --ipusher 1 > $O/r02u_vmec_rk4_$lib.json 2>> $O/r02u_err.log
  $B --workload efit_rect --ipusher 1 > $O/r02u_efit_rect_rk4_$lib.json 2>> $O/r02u_err.log
  $B --start spread > $O/r02u_vmec_spread_k2_$lib.json 2>> $O/r02u_err.log
  $B --workload efit_flux > $O/r02u_efit_flux_k2_$lib.json 2>> $O/r02u_err.log
done
for f in $O/r02u_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -5 $O/r02u_err.log
