#!/bin/bash
# round-2 batch P: group barrier as ballot + shared word + barrier.sync -- full GPU suite, synccheck + racecheck, order 3/4 throughput
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/r02p_pytest_gpu.log 2>&1
CS="timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0"
($CS --tool synccheck python -m pytest -m gpu -q tests/test_golden.py tests/test_diag_and_resort.py -k "golden or gather_and_prefetch") > $O/r02p_synccheck.log 2>&1; echo "synccheck rc=$?" > $O/r02p_summary.log
($CS --tool racecheck python -m pytest -m gpu -q tests/test_golden.py -k "golden") > $O/r02p_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02p_summary.log
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants"
$B --poly-order 4 --particles 300000 > $O/r02p_bench_vmec_k4.json 2>> $O/r02p_err.log
$B --poly-order 3 --particles 300000 > $O/r02p_bench_vmec_k3.json 2>> $O/r02p_err.log
$B --workload efit_flux --poly-order 4 --particles 300000 > $O/r02p_bench_config2_efit_flux_k4.json 2>> $O/r02p_err.log
tail -3 $O/r02p_pytest_gpu.log; cat $O/r02p_summary.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/r02p_synccheck.log $O/r02p_racecheck.log | tail -6; for f in $O/r02p_bench_*.json; do cut -c1-160 $f; done
