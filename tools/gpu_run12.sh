#!/bin/bash
# session 2 of round 2, run 12: uniform block helpers / accept copy / pos_part / Cholesky 16-byte loads
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/w_base.so $V/w_uni.so $V/w_uni_acc.so $V/w_uni_acc_pp.so $V/w_all.so $V/w_all_kb6.so > $O/r2_run12_bitcheck.log 2>&1; tail -8 $O/r2_run12_bitcheck.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run12_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run12_sweep.log
}
for B in 4096 65536; do
  for v in v_head w_base w_uni w_uni_acc w_uni_acc6 w_uni_acc_pp w_all w_all_kb6 w_all_fast v_head; do b $v $V/$v.so coop $B; done
done
QMPC_LIB=$V/w_all.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/r2_run12_coop python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > $O/r2_run12_ncu.log 2>&1
ls -la $O/r2_run12_coop.ncu-rep
