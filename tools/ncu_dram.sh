#!/bin/bash
# usage: ncu_dram.sh <lib> ... ; DRAM bytes / L2 hit rate / duration of the coop kernel at batch 16384 for each library variant
mkdir -p gpurun_out
for lib in "$@"; do
  QMPC_LIB=$PWD/scratch/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:qmpc_coop -c 1 --csv --log-file gpurun_out/dram_$lib.csv python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline > /dev/null 2>&1
  echo "== $lib" | tee -a gpurun_out/dram.log
  grep -v "^==" gpurun_out/dram_$lib.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tee -a gpurun_out/dram.log
done
