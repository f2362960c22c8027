#!/bin/bash
# round-2 batch Z: final build -- full GPU suite, smoke, config-4 bench lines with the default (cooperative) gather, ncu capture
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02z_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02z_smoke.log 2>&1
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --workload west_soledge3x > $O/r02z_bench_config4_west_rk4_strongE.json 2>> $O/r02z_err.log
$B --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02z_bench_config4_west_k2_strongE.json 2>> $O/r02z_err.log
NCU="timeout 600 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
BN="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants --workload west_soledge3x"
$NCU -o $O/r02z_west_k2_coop $BN --ipusher 2 --poly-order 2 > $O/r02z_west_k2_coop.log 2>&1
tail -n 3 $O/r02z_pytest_gpu.log; tail -n 2 $O/r02z_smoke.log
for f in $O/r02z_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], 'e2e', '%.4g'%(e.get('value') or 0), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
