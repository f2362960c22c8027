#!/bin/bash
# round-2 batch L: compute-sanitizer (memcheck, racecheck, synccheck) over small parity tests of every kernel family
mkdir -p gpurun_out
O=gpurun_out
CS="timeout 900 compute-sanitizer --error-exitcode 7 --launch-timeout 0"
T="python -m pytest -m gpu -q -x"
($CS --tool memcheck $T tests/test_golden.py tests/test_diag_and_resort.py -k "golden or gather_and_prefetch or several_streams or resort") > $O/r02l_memcheck_a.log 2>&1; echo "memcheck_a rc=$?" >> $O/r02l_summary.log
($CS --tool memcheck $T tests/test_adaptive_consumers.py tests/test_orbit_events.py tests/test_ode45.py tests/test_precomp_modes.py -k "gpu") > $O/r02l_memcheck_b.log 2>&1; echo "memcheck_b rc=$?" >> $O/r02l_summary.log
($CS --tool racecheck $T tests/test_golden.py tests/test_diag_and_resort.py -k "golden or gather_and_prefetch") > $O/r02l_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02l_summary.log
($CS --tool synccheck $T tests/test_golden.py tests/test_diag_and_resort.py -k "golden or gather_and_prefetch") > $O/r02l_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/r02l_summary.log
cat $O/r02l_summary.log; for f in memcheck_a memcheck_b racecheck synccheck; do echo == $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $O/r02l_$f.log | tail -5; done
