#!/bin/bash
# round-2 GPU run 9 (1 GPU): sanitizers, DRAM traffic record, launch list, full ncu capture, bench lines of every config
mkdir -p gpurun_out; O=gpurun_out
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 24 > $O/r2_sanitizer_memcheck.log 2>&1; tail -2 $O/r2_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize.py 24 > $O/r2_sanitizer_racecheck.log 2>&1; tail -2 $O/r2_sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 300 > $O/r2_sanitizer_memcheck_B300.log 2>&1; tail -2 $O/r2_sanitizer_memcheck_B300.log
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 > $O/r2_ncu_traffic.log 2>&1; tail -1 $O/r2_ncu_traffic.log | cut -c1-300
QMPC_COMMIT=$(cat .commit_id 2>/dev/null) timeout 200 python tools/ncu_traffic.py --batch 16384 --model convex >> $O/r2_ncu_traffic.log 2>&1; tail -1 $O/r2_ncu_traffic.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launch_list_bench_B4096.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_run9_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/r2_coop_final python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > $O/r2_run9_ncu.log 2>&1
ls -la $O/r2_coop_final.ncu-rep
for B in 1 256 65536 1048576; do timeout 200 python bench.py --batch $B --no-cpu-baseline --no-aux --no-config1 > $O/r2_bench_B$B.json 2>> $O/r2_run9_bench.err; done
timeout 200 python bench.py --batch 65536 --horizon 16 --gait mixed --no-aux --no-config1 --cpu-sample 4096 > $O/r2_bench_cfg3_N16_mixed_B65536.json 2>> $O/r2_run9_bench.err
timeout 200 python bench.py --batch 16384 --horizon 20 --model quat2 --no-aux --no-config1 --cpu-sample 4096 > $O/r2_bench_cfg4_two_contact_N20_B16384.json 2>> $O/r2_run9_bench.err
timeout 200 python bench.py --batch 16384 --model convex --no-aux --no-config1 --cpu-sample 4096 > $O/r2_bench_convex_N10_B16384.json 2>> $O/r2_run9_bench.err
timeout 200 python bench.py --batch 16384 --model convex --horizon 20 --no-aux --no-config1 --cpu-sample 2048 > $O/r2_bench_convex_N20_B16384.json 2>> $O/r2_run9_bench.err
timeout 200 python bench.py --batch 4096 --kernel phased --no-aux --no-config1 --cpu-sample 4096 > $O/r2_bench_phased_B4096.json 2>> $O/r2_run9_bench.err
timeout 300 python bench.py > $O/r2_bench_B4096.json 2>> $O/r2_run9_bench.err
for f in $O/r2_bench_*.json; do python -c "import sys,json; d=json.loads(open('$f').read()); print('$f', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['ms_per_step'], (d.get('parity') or {}).get('disagree'))"; done
