#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 30 ./tools/probes/bulk_probe 5 > $O/r2_bulk_probe_mode5.log 2>&1; echo "bulk_probe mode 5 rc=$?"; tail -1 $O/r2_bulk_probe_mode5.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run6_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run6_sweep.log
}
for B in 4096 65536; do
  for v in v_base v_inline v_bulk; do b $v $PWD/scratch/variants/$v.so coop $B; done
done
b v_inline $PWD/scratch/variants/v_inline.so coop 16384 "--model convex"
b v_base $PWD/scratch/variants/v_base.so coop 16384 "--model convex"
