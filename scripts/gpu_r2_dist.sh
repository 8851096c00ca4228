#!/bin/bash
# usage: scripts/gpu_r2_dist.sh N  -- multi-GPU parity + a sweep of the exchange pipeline's knobs (run under gpurun --gpus N)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests/test_gpu_dist.py -x -q -k "[$N]" 2>&1 | tail -5
fi
IFS=";" read -ra CFG_LIST <<< "${CFGS:-1 -1;2 -1;4 -1;4 148;4 64;8 -1}"
for cfg in "${CFG_LIST[@]}"; do
  set -- $cfg
  echo "== windows=$1 thomas_ctas=$2"
  $TR bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-parity --dist-windows $1 --dist-thomas-ctas $2 > gpurun_out/r2_dist_n${N}_w$1_c$2.log 2>&1; grep -v '^{' gpurun_out/r2_dist_n${N}_w$1_c$2.log | grep -i "error\|Traceback\|assert\|status" | head -5; grep '^{' gpurun_out/r2_dist_n${N}_w$1_c$2.log | tee gpurun_out/r2_dist_n${N}_w$1_c$2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'stage_ms', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
[ "${SKIP_FULL:-0}" = "1" ] && exit 0
echo "== full line with parity"
$TR bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/r2_bench_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'parity', d.get('parity'), 'e2e', d['e2e'])"
