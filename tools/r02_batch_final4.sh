#!/bin/bash
# round-2 final batch 4: the rebuilt library (noise options: host code and the settings struct only) -- smoke and a parity subset
mkdir -p gpurun_out
O=gpurun_out
(time timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/r02fin4_smoke.log 2>&1
(time timeout 110 python -m pytest tests/test_gpu_parity.py tests/test_mesh_options.py tests/test_orbit_events.py tests/test_rk_pusher.py -m gpu -q) > $O/r02fin4_pytest_subset.log 2>&1
grep "smoke ok" $O/r02fin4_smoke.log; tail -n 5 $O/r02fin4_pytest_subset.log | cut -c1-200
