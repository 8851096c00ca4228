#!/bin/bash
# quick GPU check: parity tests + C3 bench (device-resident only) [+ extra command]
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1
tail -4 gpurun_out/pytest.log; echo "pytest took $SECONDS s"
for w in ${WORKLOADS:-C3_channel_1024x512x512}; do
timeout 600 python bench.py --steps 20 --warmup 3 --workload $w --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | tail -1 | tee gpurun_out/bench_quick_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'ms/solve', round(d['ms_per_step'],3), 'frac', round(d['roofline']['solve']['frac'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
