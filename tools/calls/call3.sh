set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_tests3.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_tests3.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench3_n1.json 2> gpurun_out/r2_bench3_n1.err; echo "bench n1 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_bench3_n2.json 2> gpurun_out/r2_bench3_n2.err; echo "bench n2 exit $?"
grep -E "passed|failed|error|^c[135] |^fused" gpurun_out/r2_tests3.log | tail -12
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench3_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['frame_latency_ms'],4), 'launches', d['gpu_launches'], d['frame_sha'][:12], 'static4k', round(d['static_4k']['ms'],2), d['static_4k']['frame_sha'][:12], d['roofline']['stage_ms_per_frame'], d['e2e']['frame_wait_timeouts'])
    except Exception as e:
        print(n, 'failed', e)
PY
tail -5 gpurun_out/r2_bench3_n2.err
