#!/bin/bash
# run 26: compute-sanitizer synccheck / initcheck on the final kernels
mkdir -p gpurun_out; O=gpurun_out
timeout 500 compute-sanitizer --tool synccheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_synccheck.log 2>&1; tail -3 $O/r2s2_sanitizer_synccheck.log
timeout 500 compute-sanitizer --tool initcheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_initcheck.log 2>&1; tail -3 $O/r2s2_sanitizer_initcheck.log; grep -c "Uninitialized" $O/r2s2_sanitizer_initcheck.log
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_racecheck2.log 2>&1; tail -2 $O/r2s2_sanitizer_racecheck2.log
