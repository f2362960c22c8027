#!/bin/bash
# round-2 batch I2: the library as committed (template parameter renamed, no functional change) -- full GPU suite, smoke, default
# bench line; ncu captures of the RK4 kernels with the cooperative gather (EFIT rectangular, WEST strong E)
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02i2_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02i2_smoke.log 2>&1
(time timeout 900 python bench.py --no-cpu-baseline) > $O/r02i2_bench_default.json 2> $O/r02i2_bench_default.err
NCU="timeout 600 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
BN="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants --ipusher 1"
$NCU -o $O/r02i2_west_rk4_coop $BN --workload west_soledge3x > $O/r02i2_west_rk4_coop.log 2>&1
$NCU -o $O/r02i2_efit_rect_rk4_coop $BN --workload efit_rect > $O/r02i2_efit_rect_rk4_coop.log 2>&1
tail -n 3 $O/r02i2_pytest_gpu.log | head -n 2; grep "smoke ok" $O/r02i2_smoke.log; cut -c1-200 $O/r02i2_bench_default.json; ls -la $O | grep r02i2 | grep rep
