#!/bin/bash
# run 19: check-free rsqrt, small vector ops issued before the block products; gait predictor without local memory
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $V/z_rsq.so $V/z_ord.so > $O/r2_run19_bitcheck.log 2>&1; tail -4 $O/r2_run19_bitcheck.log
b() {  # name lib batch extra
  r=$(QMPC_LIB=$2 timeout 60 python bench.py --steps 5 --warmup 3 --batch $3 --no-cpu-baseline --no-aux --no-config1 $4 2>>$O/r2_run19_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$3 $4 -> $r" | tee -a $O/r2_run19_sweep.log
}
for B in 4096 65536; do
  for v in z_cvx2 z_rsq z_ord z_cvx2 z_rsq z_ord; do b $v $V/$v.so $B; done
done
timeout 200 python -m pytest tests/test_periph.py -m gpu -q -x > $O/r2_run19_periph.log 2>&1; tail -2 $O/r2_run19_periph.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config1 2>>$O/r2_run19_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v['us_per_launch'],1), round(v['frac_hbm'],3)) for k,v in d['aux_kernels'].items() if isinstance(v,dict)})" | tee -a $O/r2_run19_sweep.log
