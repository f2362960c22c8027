#!/bin/bash
# round-2 batch Q: final build -- full GPU suite, smoke, both bench arms, order 3/4
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02q_pytest_gpu.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02q_smoke.log 2>&1
(time timeout 900 python bench.py --impl reference) > $O/r02q_bench_reference.json 2> $O/r02q_bench_reference.err
(time timeout 900 python bench.py) > $O/r02q_bench_default.json 2> $O/r02q_bench_default.err
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --poly-order 4 --particles 300000 > $O/r02q_bench_vmec_k4.json 2>> $O/r02q_err.log
$B --poly-order 3 --particles 300000 > $O/r02q_bench_vmec_k3.json 2>> $O/r02q_err.log
$B --workload efit_flux --poly-order 4 --particles 300000 > $O/r02q_bench_config2_efit_flux_k4.json 2>> $O/r02q_err.log
tail -3 $O/r02q_pytest_gpu.log; tail -2 $O/r02q_smoke.log; for f in $O/r02q_bench_*.json; do cut -c1-160 $f; done
