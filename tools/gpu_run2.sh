#!/bin/bash
# round-2 GPU run 2: generic coop (quat + convex) and phased launches: tests, then A/B benches
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x > $O/r2_run2_tests.log 2>&1; tail -3 $O/r2_run2_tests.log
cp $O/parity_counts.json $O/r2_run2_parity_counts.json 2>/dev/null
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 200 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux $5 2>>$O/r2_run2_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:150])")
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run2_sweep.log
}
L=$PWD/quaternion_mpc_b200/libqmpc_b200.so
for B in 4096 65536; do
  b main $L coop $B
  b main $L phased $B
  for v in v_base v_ldgsts v_noprefetch v_acceptold; do b $v $PWD/scratch/variants/$v.so coop $B; done
  b r1kernel $PWD/scratch/lib_r1kernel.so auto $B
done
b main $L coop 16384 "--model convex"
b main $L phased 16384 "--model convex"
b main $L dense 16384 "--model convex"
b main $L coop 16384 "--model convex --horizon 20"
b main $L coop 16384 "--model quat2 --horizon 20"
b main $L phased 16384 "--model quat2 --horizon 20"
b main $L coop 65536 "--horizon 16 --gait mixed"
b main $L phased 65536 "--horizon 16 --gait mixed"
b main $L phased 256
b main $L coop 256
