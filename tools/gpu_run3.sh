#!/bin/bash
# round-2 GPU run 3: strict per-step timeouts; a smoke gate first
mkdir -p gpurun_out; O=gpurun_out
timeout 40 ./tools/probes/bulk_probe > $O/r2_bulk_probe.log 2>&1; echo "bulk_probe rc=$?"; tail -2 $O/r2_bulk_probe.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_run3_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -5 $O/r2_run3_smoke.log; exit 1; }
tail -1 $O/r2_run3_smoke.log
b() {  # name lib kernel batch extra
  r=$(QMPC_LIB=$2 timeout 90 python bench.py --steps 5 --warmup 3 --batch $4 --kernel $3 --no-cpu-baseline --no-aux $5 2>>$O/r2_run3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],4), d['kernel'][:170])" 2>/dev/null)
  echo "$1 kernel=$3 B=$4 $5 -> $r" | tee -a $O/r2_run3_sweep.log
}
L=$PWD/quaternion_mpc_b200/libqmpc_b200.so
b main $L coop 4096
b main $L phased 4096
b main $L coop 65536
b main $L phased 65536
timeout 900 python -m pytest tests -m gpu -q -s -x > $O/r2_run3_tests.log 2>&1; tail -3 $O/r2_run3_tests.log
cp $O/parity_counts.json $O/r2_run3_parity_counts.json 2>/dev/null
for B in 4096 65536; do
  for v in v_bulk v_noprefetch v_acceptold; do b $v $PWD/scratch/variants/$v.so coop $B; done
  b v_fwd168 $PWD/scratch/variants/v_fwd168.so phased $B
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
b main $L coop 1048576
