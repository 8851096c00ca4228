#!/bin/bash
# N-GPU check: distributed parity test + strong-scaling bench line
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms/solve', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'e2e', d['e2e'])"
