#!/bin/bash
# multi-GPU run: N = number of visible GPUs.  Multi-GPU C-ABI test + the driver's own torchrun bench command
N=$(nvidia-smi -L | wc -l); mkdir -p gpurun_out; O=gpurun_out
echo "GPUs: $N"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "multi_gpu" > $O/r2_multi${N}_tests.log 2>&1; tail -2 $O/r2_multi${N}_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_${N}gpu_torchrun_B4096_per_gpu.json 2> $O/r2_multi${N}_bench.err
python -c "import json; d=json.loads(open('$O/r2_bench_${N}gpu_torchrun_B4096_per_gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['path'], 'gather', d['gather'], 'parity', d['parity']['disagree'] if d['parity'] else None)"
tail -3 $O/r2_multi${N}_bench.err
