#!/bin/bash
# Development aid (run on the GPU box): the single-GPU measurement set whose outputs are summarised under profiles/.
#   bash tools/measure_n1.sh <tag>
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_n1_$tag.json 2> $out/bench_n1_$tag.err || exit 1
python bench.py --impl reference --steps 3 --warmup 3 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
# the profiled runs use direct launches (RT_GRAPH=0): ncu lists kernels launched through a graph as well, but the
# per-launch skip / count arithmetic below assumes the plain launch order
RT_GRAPH=0 python tools/frame_once.py spheres 11 1920 8 6 > $out/plain_$tag.log 2>&1 || exit 1
RT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_$tag.csv \
  python tools/frame_once.py spheres 11 1920 8 6 > $out/ncu_launches_$tag.log 2>&1
# frame_once: 6 frames x (generate, 3 x (extend, shade), tail) + 1 resolve; capture the 7 render launches of frame 4
RT_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shade|k_tail" -s 21 -c 7 -f -o $out/prof_$tag \
  python tools/frame_once.py spheres 11 1920 8 6 > $out/ncu_full_$tag.log 2>&1
# the launch list of bench.py itself (direct launches, first 700 kernels: scene build, warm-up, the timed steps and the
# end-to-end loops that follow them)
RT_GRAPH=0 python bench.py --steps 3 --warmup 3 > $out/bench_plain_$tag.log 2>&1 && \
RT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_bench_$tag.csv \
  python bench.py --steps 3 --warmup 3 > $out/ncu_launches_bench_$tag.log 2>&1
python tools/kbench.py --frames 60 --scenes c2,cornell,final,c4,c1,s14k --stats > $out/kbench_$tag.json 2> $out/kbench_$tag.err
RT_BUILD_TIMING=1 python tools/build_timing.py > $out/buildtiming_$tag.log 2>&1
python tools/run_configs.py c1 c2 c3 c4 c5 > $out/configs_n1_$tag.json 2> $out/configs_n1_$tag.err
python tools/audit_configs.py c1 c2 c3 c3n c4 c5 > $out/audit_$tag.json 2> $out/audit_$tag.err
tail -c 600 $out/bench_n1_$tag.json
