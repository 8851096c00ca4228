#!/bin/bash
# One gpurun call: smoke, GPU tests, bench (with CPU baseline), ncu launch list, ncu full set of the five
# stage kernels, tuning sweeps.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest"
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest.log 2>&1
tail -15 gpurun_out/pytest.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
tail -30 gpurun_out/launches.csv | cut -c1-250
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"r2r2_|thomas_pipe" --launch-skip 15 -c 5 \
  -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
if [ -n "$EXTRA" ]; then echo "== extra"; bash -c "$EXTRA"; fi
