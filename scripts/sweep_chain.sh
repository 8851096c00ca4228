for cc in 128 256 512; do for cs in 2 3 4 8; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --chain-cols $cc --chain-streams $cs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chain', $cc, 'streams', $cs, 'ms/solve', round(d['ms_per_step'],3))"
done; done
