#!/bin/bash
# round-2 batch Y: warp-cooperative gather with the strong-E kernels staging the whole record (45 pieces per lane, two CTAs per SM)
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest -m gpu -q -x tests/test_strong_electric_field.py tests/test_diag_and_resort.py -k "staged_gathers or gather_and_prefetch") > $O/r02y_pytest.log 2>&1
tail -n 3 $O/r02y_pytest.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
for ga in 1 2; do
  $B --gather $ga --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02y_west_k2_ga$ga.json 2>> $O/r02y_err.log
  $B --gather $ga --workload west_soledge3x > $O/r02y_west_rk4_ga$ga.json 2>> $O/r02y_err.log
done
$B --gather 2 --workload efit_rect > $O/r02y_efit_rect_k2_ga2.json 2>> $O/r02y_err.log
for f in $O/r02y_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'), int(d['counters']['pushes']), d['counters']['lost'], d['diag']['max_delta_energy'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -n 5 $O/r02y_err.log
