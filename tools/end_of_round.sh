#!/bin/bash
# end-of-session verification package on one GPU: tests, smoke, bench lines, launch list, one full ncu capture
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/eor_tests.log 2>&1; tail -2 $O/eor_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/eor_smoke.log 2>&1; tail -1 $O/eor_smoke.log
timeout 300 python bench.py > $O/eor_bench.json 2> $O/eor_bench.err; cut -c1-200 $O/eor_bench.json
timeout 300 python bench.py --impl reference > $O/eor_bench_ref.json 2>> $O/eor_bench.err; cut -c1-200 $O/eor_bench_ref.json
for B in 1 256 65536 1048576; do timeout 200 python bench.py --batch $B --no-cpu-baseline --no-aux > $O/eor_bench_B$B.json 2>> $O/eor_bench.err; done
timeout 200 python bench.py --batch 65536 --horizon 16 --gait mixed --no-cpu-baseline --no-aux > $O/eor_bench_cfg3.json 2>> $O/eor_bench.err
timeout 200 python bench.py --batch 16384 --horizon 20 --model quat2 --no-cpu-baseline --no-aux > $O/eor_bench_cfg4.json 2>> $O/eor_bench.err
for f in $O/eor_bench_B*.json $O/eor_bench_cfg*.json; do python -c "import sys,json; d=json.loads(open('$f').read()); print('$f', round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/eor_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/eor_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/eor_coop python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux > $O/eor_ncu.log 2>&1
ls -la $O/eor_coop.ncu-rep
