#!/bin/bash
# run 21: compute-sanitizer on the session-2 kernels (shared-memory overlays changed: y_k behind the dx buffer, terminal
# row hand-over, block constants)
mkdir -p gpurun_out; O=gpurun_out
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_memcheck.log 2>&1; tail -3 $O/r2s2_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py 300 > $O/r2s2_sanitizer_memcheck_B300.log 2>&1; tail -3 $O/r2s2_sanitizer_memcheck_B300.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_racecheck.log 2>&1; tail -3 $O/r2s2_sanitizer_racecheck.log
