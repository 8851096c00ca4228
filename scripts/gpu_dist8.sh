#!/bin/bash
# 8-GPU box: distributed parity (2/4/8 ranks) + strong-scaling bench lines for C3 (N = 8, 4) and C5 (N = 8)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -3
run() { # N workload tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2961$1 bench.py --gpus $1 --steps 10 --warmup 3 --workload $2 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_$3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], d['config']['workload'], 'ms/solve', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
}
run 8 C3_channel_1024x512x512 n8
run 4 C3_channel_1024x512x512 n4
run 8 C5_channel_2048x1024x1024 C5_n8
