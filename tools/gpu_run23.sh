#!/bin/bash
# run 23: L2 prefetch of the P_k rows before / after the line search
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_prod.so $V/z_pf.so > $O/r2_run23_bitcheck.log 2>&1; tail -4 $O/r2_run23_bitcheck.log
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run23_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run23_sweep.log
}
for B in 4096 65536; do
  for v in z_prod z_pf z_pf2 z_prod z_pf z_pf2; do b $v $V/$v.so $B; done
done
