#!/bin/bash
# usage: [CFGS="mode windows chunks ctas [pad_kb];..."] scripts/gpu_r2_dist.sh N
#   multi-GPU parity + a sweep of the exchange knobs (run under gpurun --gpus N)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests/test_gpu_dist.py -x -q -k "[$N]" 2>&1 | tail -5
fi
IFS=";" read -ra CFG_LIST <<< "${CFGS:-0 1 -1 -1;0 4 -1 -1;1 4 4 -1;1 2 2 -1;1 4 2 -1;1 8 4 -1}"
for cfg in "${CFG_LIST[@]}"; do
  set -- $cfg
  tag="n${N}_m$1_w$2_c$3_t$4_p${5:-a}"
  echo "== mode=$1 windows=$2 chunks=$3 thomas_ctas=$4 pad=${5:-auto}"
  $TR bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-parity --dist-mode $1 --dist-windows $2 --dist-chunks $3 --dist-thomas-ctas $4 --dist-split-pad ${5:--1} ${EXTRA:-} > gpurun_out/r2_dist_$tag.log 2>&1
  grep -v '^{' gpurun_out/r2_dist_$tag.log | grep -i "error\|Traceback\|assert\|status" | head -5
  grep '^{' gpurun_out/r2_dist_$tag.log | tee gpurun_out/r2_dist_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],4), 'stage_ms', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
[ "${SKIP_FULL:-0}" = "1" ] && exit 0
echo "== full line with parity"
$TR bench.py --gpus $N --steps 20 --warmup 3 ${EXTRA:-} > gpurun_out/r2_bench_n${N}.log 2>&1
grep '^{' gpurun_out/r2_bench_n${N}.log | tee gpurun_out/r2_bench_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'parity', d.get('parity'), 'e2e', d['e2e'])"
