#!/bin/bash
# run 20: 0.5 factor off the pivot chain, sincos in the Euler model, check-free division in the merit refresh
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_half.so $V/z_sc2.so > $O/r2_run20_bitcheck.log 2>&1; tail -4 $O/r2_run20_bitcheck.log
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run20_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run20_sweep.log
}
for B in 4096 65536; do
  for v in z_ord z_half z_sc2 z_ord z_half z_sc2; do b $v $V/$v.so $B; done
done
for v in z_ord z_sc2 z_ord z_sc2; do b $v $V/$v.so 16384 "--model convex"; done
