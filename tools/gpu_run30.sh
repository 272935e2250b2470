#!/bin/bash
# run 30: cone violation of the accepted candidate only
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/r_prod.so $V/r_viol.so > $O/r2_run30_bitcheck.log 2>&1; tail -3 $O/r2_run30_bitcheck.log
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run30_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run30_sweep.log
}
for B in 4096 65536; do
  for v in r_prod r_viol r_prod r_viol; do b $v $V/$v.so $B; done
done
