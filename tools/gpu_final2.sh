#!/bin/bash
# end-of-session verification package on one GPU (session 2 of round 2): smoke gate, full GPU tests, bench lines of every
# configuration, DRAM record, launch list, one full ncu capture
mkdir -p gpurun_out; O=gpurun_out; P=r2s2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 $O/${P}_smoke.log; exit 1; }
tail -1 $O/${P}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s -x > $O/${P}_tests.log 2>&1; tail -2 $O/${P}_tests.log
cp $O/parity_counts.json $O/${P}_parity_counts.json 2>/dev/null
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 > $O/${P}_ncu_traffic.log 2>&1; tail -1 $O/${P}_ncu_traffic.log | cut -c1-200
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 --model convex >> $O/${P}_ncu_traffic.log 2>&1
cp $O/ncu_traffic.json profiles/ncu_traffic.json
timeout 300 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err; cut -c1-200 $O/${P}_bench.json
timeout 300 python bench.py --impl reference > $O/${P}_bench_ref.json 2>> $O/${P}_bench.err; cut -c1-160 $O/${P}_bench_ref.json
for B in 1 256 65536 1048576; do timeout 200 python bench.py --batch $B --no-cpu-baseline --no-aux --no-config1 > $O/${P}_bench_B$B.json 2>> $O/${P}_bench.err; done
timeout 300 python bench.py --batch 65536 --horizon 16 --gait mixed --no-aux --no-config1 --cpu-sample 4096 > $O/${P}_bench_cfg3.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --batch 16384 --horizon 20 --model quat2 --no-aux --no-config1 --cpu-sample 4096 > $O/${P}_bench_cfg4.json 2>> $O/${P}_bench.err
timeout 200 python bench.py --batch 16384 --model convex --no-aux --no-config1 --cpu-sample 4096 > $O/${P}_bench_convex.json 2>> $O/${P}_bench.err
timeout 200 python bench.py --batch 16384 --model convex --horizon 20 --no-cpu-baseline --no-aux --no-config1 > $O/${P}_bench_convex20.json 2>> $O/${P}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launch_list_bench_B4096.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${P}_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/${P}_coop python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > $O/${P}_ncu.log 2>&1
for f in $O/${P}_bench*.json; do python -c "import sys,json; d=json.loads(open('$f').read()); print('$f', round(d['value']), round(d['e2e']['value']), round((d.get('roofline') or {}).get('frac') or 0,4), d['ms_per_step'], (d.get('parity') or {}).get('disagree'))"; done
