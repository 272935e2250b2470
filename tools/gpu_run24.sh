#!/bin/bash
# run 24: scratch split into [live x slots][trial x slots]; live part as a persisting L2 access-policy window
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/z_prod.so $V/z_split.so $V/z_persist.so > $O/r2_run24_bitcheck.log 2>&1; tail -4 $O/r2_run24_bitcheck.log
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run24_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run24_sweep.log
}
for B in 4096 65536; do
  for v in z_prod z_split z_persist z_prod z_split z_persist; do b $v $V/$v.so $B; done
done
for v in z_prod z_split z_persist; do
  QMPC_LIB=$V/$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:qmpc_coop -c 2 --csv --log-file $O/r2_run24_dram_$v.csv python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/r2_run24_dram_$v.csv")) if len(r)>14 and r[0].isdigit()]
d={}
for r in rows:
    if r[0]==rows[-1][0]: d[r[12]]=(r[14],r[13])
print("$v", d)
PY
done | tee -a $O/r2_run24_sweep.log
