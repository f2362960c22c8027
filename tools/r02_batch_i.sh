#!/bin/bash
# round-2 measurement batch I: full GPU test suite, smoke, both bench arms, ncu launch list, ncu --set full captures of the
# order-2 kernel on three workloads, the other BASELINE configs
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02i_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02i_smoke.log 2>&1
(time timeout 900 python bench.py) > $O/r02i_bench_default.json 2> $O/r02i_bench_default.err
(time timeout 900 python bench.py --impl reference) > $O/r02i_bench_reference.json 2> $O/r02i_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02i_launches_bench_default.csv \
  python bench.py --steps 2 --warmup 1 --no-variants --no-cpu-baseline > $O/r02i_launches_bench_default.log 2>&1
NCU="timeout 600 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
BN="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
$NCU -o $O/r02i_vmec_k2 $BN > $O/r02i_vmec_k2.log 2>&1
$NCU -o $O/r02i_efit_rect_k2 $BN --workload efit_rect > $O/r02i_efit_rect_k2.log 2>&1
$NCU -o $O/r02i_vmec_spread_k2 $BN --start spread > $O/r02i_vmec_spread_k2.log 2>&1
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --workload efit_rect > $O/r02i_bench_efit_rect_k2.json 2>> $O/r02i_err.log
$B --workload efit_rect --ipusher 1 > $O/r02i_bench_efit_rect_rk4.json 2>> $O/r02i_err.log
$B --workload west_soledge3x > $O/r02i_bench_config4_west_rk4_strongE.json 2>> $O/r02i_err.log
$B --workload west_soledge3x --ipusher 2 --poly-order 2 > $O/r02i_bench_config4_west_k2_strongE.json 2>> $O/r02i_err.log
$B --workload efit_flux > $O/r02i_bench_config1_efit_flux_k2.json 2>> $O/r02i_err.log
$B --workload efit_flux --poly-order 4 --particles 300000 > $O/r02i_bench_config2_efit_flux_k4.json 2>> $O/r02i_err.log
$B --poly-order 3 --particles 300000 > $O/r02i_bench_vmec_k3.json 2>> $O/r02i_err.log
$B --poly-order 4 --particles 300000 > $O/r02i_bench_vmec_k4.json 2>> $O/r02i_err.log
$B --ipusher 1 > $O/r02i_bench_vmec_rk4.json 2>> $O/r02i_err.log
$B --start spread > $O/r02i_bench_vmec_spread_k2.json 2>> $O/r02i_err.log
tail -3 $O/r02i_pytest_gpu.log; tail -2 $O/r02i_smoke.log; cat $O/r02i_bench_default.json | cut -c1-400; cat $O/r02i_bench_reference.json | cut -c1-400
ls -la $O | grep r02i
