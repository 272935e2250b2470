#!/bin/bash
# run 14: stationarity fused into the expansions pass, uniform A^T p / M^T p, cone rows in registers, phase F unroll 4
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_vec.so $V/z_all.so $V/z_statsep.so > $O/r2_run14_bitcheck.log 2>&1; tail -5 $O/r2_run14_bitcheck.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run14_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run14_sweep.log
}
for B in 4096 65536; do
  for v in x_cr z_statsep z_all z_vec x_cr z_vec; do b $v $V/$v.so coop $B; done
done
QMPC_LIB=$V/z_vec.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:qmpc_coop -c 1 -o $O/r2_run14_coop python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-aux --no-config1 > $O/r2_run14_ncu.log 2>&1
ls -la $O/r2_run14_coop.ncu-rep
