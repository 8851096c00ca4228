#!/bin/bash
# N-GPU box (N = 2 or 4): multi-GPU parity on N ranks, the C3 strong-scaling line with its parity check, the impdiff substage
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617"
mkdir -p gpurun_out
show() { grep '^{' $1 | tee ${1%.log}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', 'ms_per_step', round(d['ms_per_step'],4), 'ns/pt/solve', round(d['value'],6), 'parity', (d.get('parity') or {}).get('rel_l2'), 'e2e_ms', (d.get('e2e') or {}).get('ms_per_step'), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -5 $1 | cut -c1-300; }
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "[$N]" 2>&1 | tail -3
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; show gpurun_out/r2_bench_n$N.log C3
$TR bench.py --gpus $N --steps 10 --warmup 3 --impdiff > gpurun_out/r2_bench_C3_impdiff_n$N.log 2>&1; show gpurun_out/r2_bench_C3_impdiff_n$N.log C3_impdiff
$TR bench.py --gpus $N --steps 10 --warmup 3 --impdiff --dtdma-helmholtz > gpurun_out/r2_bench_C3_impdiff_dtdma_n$N.log 2>&1; show gpurun_out/r2_bench_C3_impdiff_dtdma_n$N.log C3_impdiff_dtdma
