#!/bin/bash
# GPU tests + cache-hint sweep of the fast transforms + C4 bench
mkdir -p gpurun_out
echo "== pytest"; SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1
tail -5 gpurun_out/pytest.log; echo "pytest took $SECONDS s"
echo "== flag sweep"
for f in 0 1 2 3 4 7; do
  echo "-- flags $f"
  timeout 300 python scripts/bench_stages.py --flags $f --no-thomas --variants 1 --kinds R2HC,HC2R 2>&1 | tail -4
done
echo "== C4"
timeout 600 python bench.py --steps 10 --warmup 3 --workload C4_duct_1024x768x768 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_C4.json
