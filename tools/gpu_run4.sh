#!/bin/bash
# round-2 GPU run 4: bulk-copy probe modes, FSM test, warm-start parity with the conditioning diagnostic, phased launch list
mkdir -p gpurun_out; O=gpurun_out
for m in 3 2 0; do timeout 30 ./tools/probes/bulk_probe $m > $O/r2_bulk_probe_mode$m.log 2>&1; echo "bulk_probe mode $m rc=$?"; tail -1 $O/r2_bulk_probe_mode$m.log; done
timeout 300 python -m pytest tests/test_periph.py tests/test_shim.py -m gpu -q -s -x > $O/r2_run4_tests_a.log 2>&1; tail -3 $O/r2_run4_tests_a.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "warm or schedule or golden or multi_gpu" > $O/r2_run4_tests_b.log 2>&1; tail -3 $O/r2_run4_tests_b.log
cp $O/parity_counts.json $O/r2_run4_parity_counts.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_phased_launches_B65536.csv python bench.py --steps 1 --warmup 1 --batch 65536 --kernel phased --no-cpu-baseline --no-aux --no-config1 > $O/r2_run4_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_phased_launches_B4096.csv python bench.py --steps 1 --warmup 1 --batch 4096 --kernel phased --no-cpu-baseline --no-aux --no-config1 >> $O/r2_run4_ncu.log 2>&1
timeout 300 python bench.py > $O/r2_run4_bench.json 2> $O/r2_run4_bench.err; cut -c1-200 $O/r2_run4_bench.json
