#!/bin/bash
# round-2 batch X: ncu captures of the config-4 kernels (WEST / SOLEDGE3X mesh, strong E): order 2 and RK4, bulk-copy gather
mkdir -p gpurun_out
O=gpurun_out
NCU="timeout 600 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
BN="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants --workload west_soledge3x"
$NCU -o $O/r02x_west_k2 $BN --ipusher 2 --poly-order 2 > $O/r02x_west_k2.log 2>&1
$NCU -o $O/r02x_west_rk4 $BN > $O/r02x_west_rk4.log 2>&1
ls -la $O | grep r02x
