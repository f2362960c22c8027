#!/bin/bash
# round-2 batch S: warp-cooperative gather (set_gather 2) -- parity of the gather modes, then throughput against modes 0 / 1
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest -m gpu -q -x tests/test_diag_and_resort.py -k "gather_and_prefetch") > $O/r02s_pytest.log 2>&1
tail -3 $O/r02s_pytest.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
for ga in 0 2; do
  $B --gather $ga > $O/r02s_vmec_k2_ga$ga.json 2>> $O/r02s_err.log
  $B --gather $ga --ipusher 1 > $O/r02s_vmec_rk4_ga$ga.json 2>> $O/r02s_err.log
done
for ga in 1 2; do
  $B --gather $ga --workload efit_rect > $O/r02s_efit_rect_k2_ga$ga.json 2>> $O/r02s_err.log
  $B --gather $ga --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02s_west_k2_ga$ga.json 2>> $O/r02s_err.log
  $B --gather $ga --workload west_soledge3x > $O/r02s_west_rk4_ga$ga.json 2>> $O/r02s_err.log
done
$B --gather 2 --start spread > $O/r02s_vmec_spread_k2_ga2.json 2>> $O/r02s_err.log
for f in $O/r02s_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -5 $O/r02s_err.log
