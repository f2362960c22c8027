#!/bin/bash
# round-2 batch F2: FINAL build of the round -- full GPU suite, smoke, both bench arms, the BASELINE-config lines whose default
# gather changed, launch list of the default bench
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02f_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02f_smoke.log 2>&1
(time timeout 900 python bench.py --impl reference) > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err
(time timeout 900 python bench.py) > $O/r02f_bench_default.json 2> $O/r02f_bench_default.err
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --workload efit_rect > $O/r02f_bench_efit_rect_k2.json 2>> $O/r02f_err.log
$B --workload efit_rect --ipusher 1 > $O/r02f_bench_efit_rect_rk4.json 2>> $O/r02f_err.log
$B --workload west_soledge3x > $O/r02f_bench_config4_west_rk4_strongE.json 2>> $O/r02f_err.log
$B --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02f_bench_config4_west_k2_strongE.json 2>> $O/r02f_err.log
$B --workload efit_flux > $O/r02f_bench_config1_efit_flux_k2.json 2>> $O/r02f_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches_bench_default.csv \
  python bench.py --steps 2 --warmup 1 --no-variants --no-cpu-baseline > $O/r02f_launches_bench_default.log 2>&1
tail -n 3 $O/r02f_pytest_gpu.log | head -n 2; grep "smoke ok" $O/r02f_smoke.log
for f in $O/r02f_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    c=d.get('clocks') or {}
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], 'e2e', '%.4g'%(e.get('value') or 0), d['ms_per_step'], c.get('sm_mhz'), c.get('reasons'), (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
