#!/bin/bash
# run 17: K_k staged through registers (ld.cg + st.shared) instead of cp.async
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_kreg.so > $O/r2_run17_bitcheck.log 2>&1; tail -3 $O/r2_run17_bitcheck.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run17_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run17_sweep.log
}
for B in 4096 65536; do
  for v in z_fin z_kreg z_fin z_kreg; do b $v $V/$v.so coop $B; done
done
