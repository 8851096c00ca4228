timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for hc in 1 4 8 16; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-chunks $hc 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks', d['e2e']['host_chunks'], 'e2e ms', round(d['e2e']['ms_per_step'],2), 'dev ms', round(d['ms_per_step'],3))"; done
timeout 600 python bench.py --steps 10 --warmup 3 --workload C4_duct_1024x768x768 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'ms/solve', round(d['ms_per_step'],3), 'frac', round(d['roofline']['solve']['frac'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
