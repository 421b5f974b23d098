#!/bin/bash
# Development aid: bench.py under different environment settings, e.g. tools/ab_env.sh RT_BVH=lbvh RT_BVH=sah
for kv in "$@"; do
  env $kv python bench.py --steps 100 --warmup 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print('$kv', round(d['ms_per_step'], 4), {a: round(b, 4) for a, b in d['roofline']['stage_ms_per_frame'].items()})"
done
