#!/bin/bash
# run 22: leg-kinematics / Raibert kernels with shared-memory staged, coalesced records
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_periph.py tests/test_shim.py -m gpu -q -x > $O/r2_run22_periph.log 2>&1; tail -2 $O/r2_run22_periph.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config1 2>>$O/r2_run22_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), {k:(round(v['us_per_launch'],1), round(v['frac_hbm'],3)) for k,v in d['aux_kernels'].items() if isinstance(v,dict)})" | tee -a $O/r2_run22_sweep.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 300 > $O/r2s2_sanitizer_memcheck_B300b.log 2>&1; tail -2 $O/r2s2_sanitizer_memcheck_B300b.log
