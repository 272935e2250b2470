#!/bin/bash
# round-2 GPU run 5: L2 discard of dead trial trajectories, 256-thread blocks, DRAM traffic, full ncu capture, bulk probe mode 4
mkdir -p gpurun_out; O=gpurun_out
timeout 30 ./tools/probes/bulk_probe 4 > $O/r2_bulk_probe_mode4.log 2>&1; echo "bulk_probe mode 4 rc=$?"; tail -1 $O/r2_bulk_probe_mode4.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_run5_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 $O/r2_run5_smoke.log; exit 1; }
tail -1 $O/r2_run5_smoke.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 90 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run5_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:170])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run5_sweep.log
}
for B in 4096 65536; do
  for v in v_discard v_nodiscard v_b256 v_f2; do b $v $PWD/scratch/variants/$v.so coop $B; done
done
b v_discard $PWD/scratch/variants/v_discard.so coop 1048576
b v_b256 $PWD/scratch/variants/v_b256.so coop 16384
b v_discard $PWD/scratch/variants/v_discard.so coop 16384
for v in v_discard v_nodiscard; do
  QMPC_LIB=$PWD/scratch/variants/$v.so timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:qmpc_coop -c 2 --csv --log-file $O/r2_dram_$v.csv python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > /dev/null 2>&1
  echo "== $v"; grep -v "^==" $O/r2_dram_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -5
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/r2_coop_run5 python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > $O/r2_run5_ncu.log 2>&1
ls -la $O/r2_coop_run5.ncu-rep
timeout 200 python -m pytest tests/test_shim.py -m gpu -q -x > $O/r2_run5_tests_shim.log 2>&1; tail -2 $O/r2_run5_tests_shim.log
