#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -3
for th in 1 2 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --workload X_2048x256x1024 --no-e2e --thomas $th 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], d['config']['workload'], 'thomas_variant', d['config']['thomas_variant'], 'ms/solve', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
