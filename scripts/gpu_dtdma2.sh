#!/bin/bash
# 2 GPUs: the implicit-diffusion Helmholtz solve on the C3 grid, transposed path vs distributed TDMA
for extra in "" "--dtdma"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --helmholtz -0.0123 $extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], d['config'].get('dtdma', False), 'ms/solve', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
