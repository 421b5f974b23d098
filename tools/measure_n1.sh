#!/bin/bash
# Development aid (run on the GPU box): the single-GPU measurement set whose outputs are summarised under profiles/.
#   bash tools/measure_n1.sh <tag>
set -u
tag=${1:-run}
out=gpurun_out
python bench.py > $out/bench_n1_$tag.json 2> $out/bench_n1_$tag.err || exit 1
python bench.py --impl reference --steps 3 --warmup 3 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
python bench.py --steps 3 --warmup 3 > $out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 3 --warmup 3 > $out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shade|k_tail" -s 21 -c 7 -f -o $out/prof_$tag \
  python bench.py --steps 3 --warmup 3 > $out/ncu_full_$tag.log 2>&1
python tools/run_configs.py c1 c2 c3 c4 c5 > $out/configs_n1_$tag.json 2> $out/configs_n1_$tag.err
tail -c 600 $out/bench_n1_$tag.json
