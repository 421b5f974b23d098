#!/bin/bash
# Development aid: bench.py against the kernel variants in build/variants (RT_B200_LIB override).
for lib in "" $(ls build/variants/librt_*.so 2>/dev/null); do
  RT_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 100 --warmup 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print('${lib:-default}'.split('/')[-1], round(d['ms_per_step'], 4), {a: round(b, 4) for a, b in d['roofline']['stage_ms_per_frame'].items()})"
done
