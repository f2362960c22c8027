#!/bin/bash
# round-2 batch H2: RK4 kernels at two CTAs per SM (255 registers, no spills; variant library rk2) against the shipped three
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/gorilla_b200/lib/libgorilla_b200_rk2.so
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --ipusher 1"
for lib in main rk2; do
  if [ $lib = rk2 ]; then export GORILLA_B200_LIB=$V; else unset GORILLA_B200_LIB; fi
  $B --gather 0 > $O/r02h_vmec_rk4_ga0_$lib.json 2>> $O/r02h_err.log
  $B --gather 0 --workload efit_flux > $O/r02h_efit_flux_rk4_ga0_$lib.json 2>> $O/r02h_err.log
done
for f in $O/r02h_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'), int(d['counters']['pushes']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -n 3 $O/r02h_err.log
