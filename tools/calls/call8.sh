set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests8.log 2>&1; tail -3 gpurun_out/r2_tests8.log
for wb in 2 3 4 5 6; do RT_WAVE_BOUNCES=$wb timeout 300 python tools/kbench.py --frames 40 --scenes c2,cornell 2>> gpurun_out/r2_kbench8.err | sed "s/\"lib\": \"default\"/\"lib\": \"wave$wb\"/" >> gpurun_out/r2_kbench8.json; done
cut -c1-260 gpurun_out/r2_kbench8.json
timeout 900 python tools/ref_gpu_frame.py spheres 11 1920 8 1 > gpurun_out/r2_refgpu4.json 2> gpurun_out/r2_refgpu4.err; cat gpurun_out/r2_refgpu4.json
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench8_n1.json 2> gpurun_out/r2_bench8_n1.err; echo "bench n1 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench8_n2.json 2> gpurun_out/r2_bench8_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench8_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['frame_latency_ms'],4), 'launches', d['gpu_launches'], d['frame_sha'][:12], 'static4k', round(d['static_4k']['ms'],2), d['static_4k']['frame_sha'][:12], d['roofline']['stage_ms_per_frame'], d['e2e']['frame_wait_timeouts'])
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/r2_bench8_n2.err
