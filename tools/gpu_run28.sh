#!/bin/bash
# last check of HEAD: smoke + full GPU tests + one bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2s2_last_tests.log 2>&1; tail -2 $O/r2s2_last_tests.log
timeout 300 python bench.py > $O/r2s2_last_bench.json 2> $O/r2s2_last_bench.err; python -c "import json; d=json.loads(open('$O/r2s2_last_bench.json').read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['parity']['disagree'], d['gpu_launches'], d['clocks'])"
