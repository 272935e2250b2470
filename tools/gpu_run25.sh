#!/bin/bash
# run 25: unroll / block-size knobs re-measured on the session-2 kernel (21 % fewer instructions than when they were set)
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run25_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run25_sweep.log
}
for B in 4096 65536; do
  for v in p_base p_foot2 p_blk1 p_f2 p_f6 p_b256 p_kb3 p_base; do b $v $V/$v.so $B; done
done
