#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run11_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run11_sweep.log
}
for B in 4096 65536; do
  for v in v_base v_blk3 v_cholfwd v_both v_base; do b $v $PWD/scratch/variants/$v.so coop $B; done
done
b v_both $PWD/scratch/variants/v_both.so coop 16384 "--model convex"
b v_base $PWD/scratch/variants/v_base.so coop 16384 "--model convex"
