#!/bin/bash
# round-2 batch N: after the __syncwarp() fix of the group barriers -- full GPU suite, synccheck again, order 3/4 throughput, default bench
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02n_pytest_gpu.log 2>&1
(timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0 --tool synccheck python -m pytest -m gpu -q -x tests/test_golden.py tests/test_diag_and_resort.py -k "golden or gather_and_prefetch") > $O/r02n_synccheck.log 2>&1; echo "synccheck rc=$?" > $O/r02n_summary.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --poly-order 4 --particles 300000 > $O/r02n_bench_vmec_k4.json 2>> $O/r02n_err.log
$B --poly-order 3 --particles 300000 > $O/r02n_bench_vmec_k3.json 2>> $O/r02n_err.log
(time timeout 900 python bench.py) > $O/r02n_bench_default.json 2> $O/r02n_bench_default.err
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02n_smoke.log 2>&1
tail -3 $O/r02n_pytest_gpu.log; cat $O/r02n_summary.log; grep -E "ERROR SUMMARY|passed|failed" $O/r02n_synccheck.log | tail -3; for f in $O/r02n_bench_*.json; do cut -c1-160 $f; done; tail -2 $O/r02n_smoke.log
