#!/bin/bash
# Development aid (run on an N-GPU box): bench.py and the C5 configuration on N GPUs, one process per GPU.
#   bash tools/measure_multi.sh <N> <tag>
set -u
n=$1
tag=${2:-run}
out=gpurun_out
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"
$run bench.py --gpus $n > $out/bench_n${n}_$tag.json 2> $out/bench_n${n}_$tag.err
$run tools/run_configs.py c5 > $out/c5_n${n}_$tag.json 2> $out/c5_n${n}_$tag.err
tail -c 400 $out/bench_n${n}_$tag.json; tail -c 400 $out/c5_n${n}_$tag.json
