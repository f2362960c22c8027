#!/bin/bash
# round-2 batch A2: PHI = 1 kernels (mesh with an electrostatic potential, eps_Phi = -1e-7) on the big EFIT mesh: vector loads,
# bulk copies, cooperative gather with the whole record staged; gather-mode parity test again (covers PHI = 1, small mesh)
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest -m gpu -q -x tests/test_strong_electric_field.py tests/test_diag_and_resort.py -k "staged_gathers or gather_and_prefetch") > $O/r02a2_pytest.log 2>&1
tail -n 3 $O/r02a2_pytest.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --workload efit_rect --eps-phi=-1e-7"
for ga in 0 1 2; do
  $B --gather $ga > $O/r02a2_efit_rect_phi_k2_ga$ga.json 2>> $O/r02a2_err.log
done
for ga in 1 2; do
  $B --gather $ga --ipusher 1 > $O/r02a2_efit_rect_phi_rk4_ga$ga.json 2>> $O/r02a2_err.log
done
for f in $O/r02a2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'), int(d['counters']['pushes']), d['counters']['lost'], d['diag']['max_delta_energy'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -n 5 $O/r02a2_err.log
