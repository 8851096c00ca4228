#!/bin/bash
# One gpurun call: GPU tests, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest"
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest.log 2>&1
tail -40 gpurun_out/pytest.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$EXTRA" ]; then echo "== extra"; bash -c "$EXTRA"; fi
