#!/bin/bash
# bench sweeps over tuning knobs; prints ms_per_step per configuration
for args in "$@"; do
  echo -n "$args => "
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
