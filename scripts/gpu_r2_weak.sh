#!/bin/bash
# weak scaling of the C5 operator: 2048 x 1024 x (128 N) on N GPUs (per-GPU slab = C5's at 8 GPUs); Poisson solve and impdiff substage
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29619"; fi
for mode in "" "--impdiff"; do
  tag="r2_weak_C5_n${N}${mode:+_impdiff}"
  $TR bench.py --gpus $N --steps 10 --warmup 3 --workload C5_channel_2048x1024x1024 --weak-nz 128 --no-e2e --no-cpu-baseline --no-parity $mode > gpurun_out/$tag.log 2>&1
  grep '^{' gpurun_out/$tag.log | tee gpurun_out/$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag', 'ms_per_step', round(d['ms_per_step'],4), 'ns/pt/solve', round(d['value'],6), d['scaling'], {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -5 gpurun_out/$tag.log | cut -c1-300
done
