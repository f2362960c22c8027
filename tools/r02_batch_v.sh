#!/bin/bash
# round-2 batch V: build with the warp-cooperative gather as the default on the big magnetic-only meshes -- full GPU suite,
# smoke, both bench arms, the workloads whose default gather changed, launch list, compute-sanitizer on the new kernels
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02v_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02v_smoke.log 2>&1
(time timeout 900 python bench.py --impl reference) > $O/r02v_bench_reference.json 2> $O/r02v_bench_reference.err
(time timeout 900 python bench.py) > $O/r02v_bench_default.json 2> $O/r02v_bench_default.err
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --workload efit_rect > $O/r02v_bench_efit_rect_k2.json 2>> $O/r02v_err.log
$B --workload efit_rect --ipusher 1 > $O/r02v_bench_efit_rect_rk4.json 2>> $O/r02v_err.log
$B --workload west_soledge3x > $O/r02v_bench_config4_west_rk4_strongE.json 2>> $O/r02v_err.log
$B --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02v_bench_config4_west_k2_strongE.json 2>> $O/r02v_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02v_launches_bench_efit_rect.csv \
  python bench.py --steps 2 --warmup 1 --no-variants --no-cpu-baseline --workload efit_rect > $O/r02v_launches_bench_efit_rect.log 2>&1
T="python -m pytest -m gpu -q tests/test_diag_and_resort.py -k gather_and_prefetch"
(timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0 --tool memcheck $T) > $O/r02v_memcheck.log 2>&1; echo "memcheck rc=$?" > $O/r02v_summary.log
(timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0 --tool racecheck $T) > $O/r02v_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02v_summary.log
(timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0 --tool synccheck --num-cuda-barriers 400000 $T) > $O/r02v_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/r02v_summary.log
tail -n 3 $O/r02v_pytest_gpu.log; tail -n 2 $O/r02v_smoke.log; cat $O/r02v_summary.log
for f in $O/r02v_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], 'e2e', '%.4g'%(e.get('value') or 0), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
for t in memcheck racecheck synccheck; do grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/r02v_$t.log | tail -n 2; done
