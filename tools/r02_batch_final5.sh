#!/bin/bash
# round-2 final batch 5: the rest of the GPU suite on the rebuilt library (all files but the four of batch 4 and the full-size parity file)
mkdir -p gpurun_out
O=gpurun_out
(time timeout 115 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_parity.py --ignore=tests/test_mesh_options.py --ignore=tests/test_orbit_events.py --ignore=tests/test_rk_pusher.py --ignore=tests/test_full_size_parity.py) > $O/r02fin5_pytest_rest.log 2>&1
tail -n 5 $O/r02fin5_pytest_rest.log | cut -c1-200
