#!/bin/bash
# run 18: product library = z_cvx kernels + chunked host pipeline: bit-identity, host-path tests, e2e A/B
mkdir -p gpurun_out; O=gpurun_out; V=$PWD/scratch/variants
timeout 300 python tools/gpu_bitcheck.py $V/v_head.so $PWD/quaternion_mpc_b200/libqmpc_b200.so > $O/r2_run18_bitcheck.log 2>&1; tail -3 $O/r2_run18_bitcheck.log
timeout 600 python -m pytest tests -m gpu -q -x -k "chunked or host_and_device or multi_gpu or shim or convex" > $O/r2_run18_tests.log 2>&1; tail -3 $O/r2_run18_tests.log
b() {  # name batch extra
  r=$(timeout 100 python bench.py --steps 10 --warmup 3 --batch $2 --no-cpu-baseline --no-aux --no-config1 $3 2>>$O/r2_run18_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4))" 2>/dev/null)
  echo "$1 B=$2 $3 -> $r" | tee -a $O/r2_run18_sweep.log
}
for B in 4096 65536; do
  b chunks $B ""; b nochunks $B "--host-chunks 1"; b chunks $B ""; b nochunks $B "--host-chunks 1"
done
b convex 16384 "--model convex"; QMPC_LIB=$V/z_fin.so b convex_zfin 16384 "--model convex"
b convex20 16384 "--model convex --horizon 20"
