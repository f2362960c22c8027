#!/bin/bash
# round-2 batch M: precomputed-coefficient modes on a right-handed mesh (analytic tokamak, 768 000 tetrahedra), then batch L
mkdir -p gpurun_out
O=gpurun_out
B="timeout 400 python bench.py --workload analytic --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B > $O/r02m_bench_analytic_k2.json 2>> $O/r02m_err.log
$B --i-precomp 1 > $O/r02m_bench_analytic_k2_precomp1.json 2>> $O/r02m_err.log
$B --i-precomp 2 > $O/r02m_bench_analytic_k2_precomp2.json 2>> $O/r02m_err.log
$B --poly-order 4 --particles 300000 > $O/r02m_bench_analytic_k4.json 2>> $O/r02m_err.log
$B --poly-order 4 --particles 300000 --i-precomp 1 > $O/r02m_bench_analytic_k4_precomp1.json 2>> $O/r02m_err.log
$B --poly-order 3 --particles 300000 > $O/r02m_bench_analytic_k3.json 2>> $O/r02m_err.log
$B --poly-order 3 --particles 300000 --i-precomp 1 > $O/r02m_bench_analytic_k3_precomp1.json 2>> $O/r02m_err.log
$B --ipusher 1 > $O/r02m_bench_analytic_rk4.json 2>> $O/r02m_err.log
$B --ipusher 1 --newton-precalc > $O/r02m_bench_analytic_rk4_newton_precalc.json 2>> $O/r02m_err.log
(time timeout 900 python tools/config2_conservation.py 100000 300 4) > $O/r02m_config2_conservation.json 2> $O/r02m_config2.err
(time timeout 900 python tools/config3_loss.py 1000000 500) > $O/r02m_config3_full_run.json 2> $O/r02m_config3.err
for f in $O/r02m_bench_*.json; do echo $f; cut -c1-120 $f; done; tail -5 $O/r02m_err.log; cut -c1-600 $O/r02m_config2_conservation.json; tail -3 $O/r02m_config2.err; cut -c1-600 $O/r02m_config3_full_run.json
bash tools/r02_batch_l.sh
