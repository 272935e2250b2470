#!/bin/bash
# round-2 GPU run 1: full GPU test suite with parity counts, smoke, bench line, DMMA probe (+ncu pipe metrics)
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/r2_gpus.txt; nproc >> $O/r2_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -s -x > $O/r2_run1_tests.log 2>&1; tail -3 $O/r2_run1_tests.log
cp $O/parity_counts.json $O/r2_run1_parity_counts.json 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_run1_smoke.log 2>&1; tail -1 $O/r2_run1_smoke.log
timeout 400 python bench.py > $O/r2_run1_bench.json 2> $O/r2_run1_bench.err; cut -c1-300 $O/r2_run1_bench.json; tail -3 $O/r2_run1_bench.err
timeout 120 ./tools/probes/dmma_probe > $O/r2_dmma_probe.json 2>&1; cat $O/r2_dmma_probe.json
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum --clock-control none --csv --log-file $O/r2_dmma_probe_ncu.csv ./tools/probes/dmma_probe > /dev/null 2>&1
tail -n +1 $O/r2_dmma_probe_ncu.csv | cut -c1-220 | tail -40
