#!/bin/bash
# in-tree library, spread vs packed partial waves at several batch sizes
mkdir -p gpurun_out
for B in 1 64 256 512 1024 1184 1500; do
  for ns in 0 1; do
    if [ "$ns" = "1" ]; then export QMPC_COOP_NO_SPREAD=1; else unset QMPC_COOP_NO_SPREAD; fi
    r=$(timeout 100 python bench.py --steps 5 --warmup 3 --batch $B --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'])")
    echo "no_spread=$ns B=$B -> $r" | tee -a gpurun_out/spread.log
  done
done
