#!/bin/bash
# round-2 GPU run 8: validation of the new default (inline roll-out, discard, prefetch): full tests, sanitizers, bench lines
mkdir -p gpurun_out; O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_run8_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 $O/r2_run8_smoke.log; exit 1; }
tail -1 $O/r2_run8_smoke.log
timeout 900 python -m pytest tests -m gpu -q -s -x > $O/r2_run8_tests.log 2>&1; tail -3 $O/r2_run8_tests.log
cp $O/parity_counts.json $O/r2_run8_parity_counts.json 2>/dev/null
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 24 > $O/r2_sanitizer_memcheck.log 2>&1; tail -3 $O/r2_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize.py 24 > $O/r2_sanitizer_racecheck.log 2>&1; tail -3 $O/r2_sanitizer_racecheck.log
timeout 300 python bench.py > $O/r2_run8_bench.json 2> $O/r2_run8_bench.err; cut -c1-250 $O/r2_run8_bench.json
timeout 300 python bench.py --impl reference > $O/r2_run8_bench_ref.json 2>> $O/r2_run8_bench.err; cut -c1-250 $O/r2_run8_bench_ref.json
