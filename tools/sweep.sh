#!/bin/bash
# usage: sweep.sh  "<lib>:<WIDE or ->" ...   ; runs bench at two batch sizes for each variant
mkdir -p gpurun_out
for v in "$@"; do
  lib=${v%%:*}; wide=${v#*:}
  for B in 65536 4096; do
    if [ "$wide" = "-" ]; then unset QMPC_COOP_WIDE; else export QMPC_COOP_WIDE=$wide; fi
    r=$(QMPC_LIB=$PWD/scratch/$lib timeout 100 python bench.py --steps 5 --warmup 3 --batch $B --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['config'].get('kernel',''))")
    echo "$lib wide=$wide B=$B -> $r" | tee -a gpurun_out/sweep.log
  done
done
