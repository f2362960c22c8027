#!/bin/bash
# round-2 batch W: L2 prefetch of the sub-records that stay per-lane loads (Phi, strong E) next to the staged gathers
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest -m gpu -q -x tests/test_diag_and_resort.py -k "gather_and_prefetch") > $O/r02w_pytest.log 2>&1
tail -n 2 $O/r02w_pytest.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
for pf in 0 1; do
  $B --gather 1 --prefetch $pf --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02w_west_k2_ga1_pf$pf.json 2>> $O/r02w_err.log
  $B --gather 1 --prefetch $pf --workload west_soledge3x > $O/r02w_west_rk4_ga1_pf$pf.json 2>> $O/r02w_err.log
  $B --gather 2 --prefetch $pf --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02w_west_k2_ga2_pf$pf.json 2>> $O/r02w_err.log
done
$B --gather 2 --prefetch 1 --workload west_soledge3x > $O/r02w_west_rk4_ga2_pf1.json 2>> $O/r02w_err.log
for f in $O/r02w_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'), int(d['counters']['pushes']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -n 5 $O/r02w_err.log
