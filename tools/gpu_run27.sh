#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 500 compute-sanitizer --tool initcheck python tools/sanitize.py 24 > $O/r2s2_sanitizer_initcheck.log 2>&1; tail -3 $O/r2s2_sanitizer_initcheck.log; grep "at .*\.cuh\|at .*\.cu:" $O/r2s2_sanitizer_initcheck.log | sed -E 's/.*at //' | sort | uniq -c | sort -rn | head
