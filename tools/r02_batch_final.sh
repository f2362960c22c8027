#!/bin/bash
# round-2 final batch: the library as committed at the end of the round (host mesh options, full-orbit events, struct-size check)
# -- full GPU suite (no -x: every failure is seen), smoke, default bench line
mkdir -p gpurun_out
O=gpurun_out
(time timeout 330 python -m pytest tests -m gpu -q) > $O/r02fin_pytest_gpu.log 2>&1
(time timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02fin_smoke.log 2>&1
(time timeout 120 python bench.py --no-cpu-baseline) > $O/r02fin_bench_default.json 2> $O/r02fin_bench_default.err
tail -n 12 $O/r02fin_pytest_gpu.log | cut -c1-300; grep "smoke ok" $O/r02fin_smoke.log; cut -c1-400 $O/r02fin_bench_default.json
