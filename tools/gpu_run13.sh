#!/bin/bash
# run 13: lxx folded into phase C, Cholesky quotient correction with r0, check-free divisions, y_k in shared memory
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/x_all.so $V/x_rdg.so $V/x_ieee.so $V/x_cr.so > $O/r2_run13_bitcheck.log 2>&1; tail -8 $O/r2_run13_bitcheck.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux --no-config1 $5 2>>$O/r2_run13_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:100])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run13_sweep.log
}
for B in 4096 65536; do
  for v in w_all x_all x_rmw x_rdg x_ieee x_dxg x_cr x_f4 x_f12 w_all x_all; do b $v $V/$v.so coop $B; done
done
for v in w_all x_all; do b $v $V/$v.so coop 65536 "--horizon 16 --gait mixed"; b $v $V/$v.so coop 16384 "--horizon 20 --model quat2"; b $v $V/$v.so coop 16384 "--model convex"; done
