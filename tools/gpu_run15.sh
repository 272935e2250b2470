#!/bin/bash
# run 15: explicit-fma uniform A^T p / M^T p, terminal row through shared memory, accept rows over 16 lanes + prefetch, CR LDS.128
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_vec2.so $V/z_term.so $V/z_acc16.so $V/z_new.so > $O/r2_run15_bitcheck.log 2>&1; tail -6 $O/r2_run15_bitcheck.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run15_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run15_sweep.log
}
for B in 4096 65536; do
  for v in z_statsep z_vec2 z_term z_acc16 z_new z_statsep z_new; do b $v $V/$v.so coop $B; done
done
