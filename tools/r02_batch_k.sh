#!/bin/bash
# round-2 batch K: full GPU suite after the EXT = 5 kernels, precomputed-coefficient / ODE45 / combined-adaptive benches, rebin capture
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02k_pytest_gpu.log 2>&1
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B > $O/r02k_bench_default_quick.json 2>> $O/r02k_err.log
$B --workload west_soledge3x > $O/r02k_bench_config4_west_rk4_strongE.json 2>> $O/r02k_err.log
$B --poly-order 4 --particles 300000 --i-precomp 1 > $O/r02k_bench_vmec_k4_precomp1.json 2>> $O/r02k_err.log
$B --poly-order 3 --particles 300000 --i-precomp 1 > $O/r02k_bench_vmec_k3_precomp1.json 2>> $O/r02k_err.log
$B --i-precomp 1 > $O/r02k_bench_vmec_k2_precomp1.json 2>> $O/r02k_err.log
$B --i-precomp 2 > $O/r02k_bench_vmec_k2_precomp2.json 2>> $O/r02k_err.log
$B --ipusher 1 --newton-precalc > $O/r02k_bench_vmec_rk4_newton_precalc.json 2>> $O/r02k_err.log
$B --ipusher 1 --ode45 --particles 300000 > $O/r02k_bench_vmec_rk_ode45.json 2>> $O/r02k_err.log
$B --adaptive 1e-7 --time-tracing 2 --particles 200000 > $O/r02k_bench_vmec_k2_adaptive_hamiltonian.json 2>> $O/r02k_err.log
$B --adaptive 1e-10 --time-tracing 2 --poly-order 4 --particles 100000 > $O/r02k_bench_vmec_k4_adaptive_hamiltonian.json 2>> $O/r02k_err.log
NCU="timeout 900 ncu --set full --import-source on --clock-control none -k regex:orbit_kernel -s 2 -c 1 -f"
$NCU -o $O/r02k_vmec_k4_rebin1 python bench.py --poly-order 4 --particles 300000 --rebin 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants > $O/r02k_vmec_k4_rebin1.log 2>&1
tail -3 $O/r02k_pytest_gpu.log; for f in $O/r02k_bench_*.json; do echo $f; cut -c1-120 $f; done; tail -5 $O/r02k_err.log
