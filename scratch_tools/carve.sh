#!/bin/bash
# usage: carve.sh <lib> "<wide>:<carveout %>" ... ; bench at two batch sizes per (layout, carveout) pair
mkdir -p gpurun_out
lib=$1; shift
for v in "$@"; do
  wide=${v%%:*}; carve=${v#*:}
  for B in 65536 4096; do
    if [ "$wide" = "-" ]; then unset QMPC_COOP_WIDE; else export QMPC_COOP_WIDE=$wide; fi
    if [ "$carve" = "-" ]; then unset QMPC_COOP_CARVEOUT; else export QMPC_COOP_CARVEOUT=$carve; fi
    r=$(QMPC_LIB=$PWD/scratch/$lib timeout 100 python bench.py --steps 5 --warmup 3 --batch $B --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['config'].get('kernel','')[60:])")
    echo "$lib wide=$wide carve=$carve B=$B -> $r" | tee -a gpurun_out/carve.log
  done
done
