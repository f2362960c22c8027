#!/bin/bash
# round-2 batch T: ncu captures of the cooperative-gather kernel (VMEC headline, EFIT rectangular), remaining mode comparisons
mkdir -p gpurun_out
O=gpurun_out
NCU="timeout 600 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
BN="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants --gather 2"
$NCU -o $O/r02t_vmec_k2_ga2 $BN > $O/r02t_vmec_k2_ga2.log 2>&1
$NCU -o $O/r02t_efit_rect_k2_ga2 $BN --workload efit_rect > $O/r02t_efit_rect_k2_ga2.log 2>&1
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
for ga in 1 2; do
  $B --gather $ga --workload efit_rect --ipusher 1 > $O/r02t_efit_rect_rk4_ga$ga.json 2>> $O/r02t_err.log
done
for ga in 0 2; do
  $B --gather $ga --workload efit_flux > $O/r02t_efit_flux_k2_ga$ga.json 2>> $O/r02t_err.log
done
$B --gather 0 --workload efit_rect > $O/r02t_efit_rect_k2_ga0.json 2>> $O/r02t_err.log
for f in $O/r02t_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['roofline'].get('kernel'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -5 $O/r02t_err.log; ls -la $O | grep r02t
