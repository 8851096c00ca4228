timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for zm in 0 1; do for w in C3_channel_1024x512x512 C2_tgv_512x512x512 C4_duct_1024x768x768; do
timeout 300 python bench.py --steps 20 --warmup 3 --workload $w --no-cpu-baseline --no-e2e --zmajor $zm 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('zmajor $zm', d['config']['workload'], 'ms/solve', round(d['ms_per_step'],3), 'frac', round(d['roofline']['solve']['frac'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done; done
