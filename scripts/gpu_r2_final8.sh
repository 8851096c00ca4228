#!/bin/bash
# 8-GPU box, one call: multi-GPU parity on 8 ranks, the C3 strong-scaling line with its parity check, and the C5 lines
# (Poisson; one RK substage of an is_impdiff run with the Helmholtz solves on the transposes / on the distributed TDMA)
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618"
mkdir -p gpurun_out
show() { grep '^{' $1 | tee ${1%.log}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', 'ms_per_step', round(d['ms_per_step'],4), 'ns/pt/solve', round(d['value'],6), 'parity', (d.get('parity') or {}).get('rel_l2'), 'e2e_ms', (d.get('e2e') or {}).get('ms_per_step'), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -5 $1 | cut -c1-300; }
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "[8]" 2>&1 | tail -3
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.log 2>&1; show gpurun_out/r2_bench_n8.log C3
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload C5_channel_2048x1024x1024 --no-e2e > gpurun_out/r2_bench_C5_n8.log 2>&1; show gpurun_out/r2_bench_C5_n8.log C5_poisson
$TR bench.py --gpus $N --steps 6 --warmup 3 --workload C5_channel_2048x1024x1024 --impdiff > gpurun_out/r2_bench_C5_impdiff_n8.log 2>&1; show gpurun_out/r2_bench_C5_impdiff_n8.log C5_impdiff
$TR bench.py --gpus $N --steps 6 --warmup 3 --workload C5_channel_2048x1024x1024 --impdiff --dtdma-helmholtz > gpurun_out/r2_bench_C5_impdiff_dtdma_n8.log 2>&1; show gpurun_out/r2_bench_C5_impdiff_dtdma_n8.log C5_impdiff_dtdma
$TR bench.py --gpus $N --steps 10 --warmup 3 --impdiff --dtdma-helmholtz > gpurun_out/r2_bench_C3_impdiff_dtdma_n8.log 2>&1; show gpurun_out/r2_bench_C3_impdiff_dtdma_n8.log C3_impdiff_dtdma
$TR bench.py --gpus $N --steps 10 --warmup 3 --impdiff > gpurun_out/r2_bench_C3_impdiff_n8.log 2>&1; show gpurun_out/r2_bench_C3_impdiff_n8.log C3_impdiff
