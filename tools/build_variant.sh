#!/bin/bash
# Development aid: build/variants/librt_<name>.so = the CUDA library with extra -D flags on rt_kernels.cu
# (A/B measurements: tools/kbench.py, tools/ab.sh pick the variants up through RT_B200_LIB).
#   tools/build_variant.sh <name> [-DRT_SHADE_SORT=1 ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/real-time-ray-tracing-engine_b200/csrc
out=$root/build/variants
mkdir -p $out/obj_$name
flags="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
nvcc $flags "$@" -Xptxas -v -c $src/rt_kernels.cu -o $out/obj_$name/rt_kernels.o 2> $out/obj_$name/ptxas.log
for f in rt_exact rt_scene rt_api rt_frame; do
  [ $src/$f.o -nt $src/$f.cu ] || make -C $src $f.o > /dev/null
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/librt_$name.so $out/obj_$name/rt_kernels.o $src/rt_exact.o $src/rt_scene.o $src/rt_api.o $src/rt_frame.o
grep -E "Compiling entry|Used|spill" $out/obj_$name/ptxas.log | sed 's/ptxas info    : //' | grep -A2 -E "k_extendILb0|k_tailILb0|k_shadeILb" | grep -E "Used|spill" | paste - - | awk '{print "  " $0}' | cut -c1-200
echo "built $out/librt_$name.so"
