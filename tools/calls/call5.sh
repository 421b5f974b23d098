set -x
mkdir -p gpurun_out
timeout 600 python tools/ref_gpu_frame.py spheres 11 1920 8 3 > gpurun_out/r2_refgpu.json 2> gpurun_out/r2_refgpu.err; echo "refgpu exit $?"
cat gpurun_out/r2_refgpu.json; tail -3 gpurun_out/r2_refgpu.err
RT_BUILD_TIMING=1 timeout 300 python tools/frame_once.py spheres_textured 500 1920 8 1 > gpurun_out/r2_buildtiming2.log 2>&1
cat gpurun_out/r2_buildtiming2.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['frame_latency_ms'],4), 'launches', d['gpu_launches'])
print(json.dumps(d['gpu_reference'])[:900])
PY
