set -x
mkdir -p gpurun_out
N=$1
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $run bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench12_n$N.json 2> gpurun_out/r2_bench12_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench12_n$N.json').read().strip().splitlines()[-1])
print($N, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['frame_latency_ms'],4), 'launches', d['gpu_launches'], d['frame_sha'][:12], 'static4k', round(d['static_4k']['ms'],2), d['static_4k']['frame_sha'][:12], d['e2e']['frame_wait_timeouts'])
PY
tail -3 gpurun_out/r2_bench12_n$N.err
