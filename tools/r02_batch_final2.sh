#!/bin/bash
# round-2 final batch 2: memcheck of the new full-orbit event path, throughput of the event kernels with it
mkdir -p gpurun_out
O=gpurun_out
(time timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_orbit_events.py -m gpu -q -k "full_orbit and 2-2") > $O/r02fin_memcheck_full_orbit.log 2>&1
echo "memcheck rc=$?" >> $O/r02fin_memcheck_full_orbit.log
(time timeout 110 python tools/full_orbit_rate.py 200000 50) > $O/r02fin_full_orbit_rate.json 2> $O/r02fin_full_orbit_rate.err
tail -n 8 $O/r02fin_memcheck_full_orbit.log | cut -c1-200; cut -c1-1200 $O/r02fin_full_orbit_rate.json; tail -n 3 $O/r02fin_full_orbit_rate.err
