#!/bin/bash
# end-of-session verification package on one GPU: smoke gate, full GPU tests, bench lines, launch list, DRAM record
mkdir -p gpurun_out; O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 $O/r2_final_smoke.log; exit 1; }
tail -1 $O/r2_final_smoke.log
timeout 900 python -m pytest tests -m gpu -q -s -x > $O/r2_final_tests.log 2>&1; tail -2 $O/r2_final_tests.log
cp $O/parity_counts.json $O/r2_final_parity_counts.json 2>/dev/null
timeout 300 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; cut -c1-200 $O/r2_final_bench.json
timeout 300 python bench.py --impl reference > $O/r2_final_bench_ref.json 2>> $O/r2_final_bench.err; cut -c1-160 $O/r2_final_bench_ref.json
for B in 256 65536; do timeout 200 python bench.py --batch $B --no-cpu-baseline --no-aux --no-config1 > $O/r2_final_bench_B$B.json 2>> $O/r2_final_bench.err; done
timeout 200 python bench.py --batch 16384 --model convex --no-aux --no-config1 --cpu-sample 4096 > $O/r2_final_bench_convex.json 2>> $O/r2_final_bench.err
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 > $O/r2_final_ncu_traffic.log 2>&1; tail -1 $O/r2_final_ncu_traffic.log | cut -c1-200
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 --model convex >> $O/r2_final_ncu_traffic.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final_launch_list_bench_B4096.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_final_b.log 2>&1
for f in $O/r2_final_bench*.json; do python -c "import sys,json; d=json.loads(open('$f').read()); print('$f', round(d['value']), round(d['e2e']['value']), round((d.get('roofline') or {}).get('frac') or 0,4), d['ms_per_step'], (d.get('parity') or {}).get('disagree'))"; done
