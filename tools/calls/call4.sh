set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests4.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_tests4.log
tail -4 gpurun_out/r2_tests4.log
for mode in auto lbvh ploc sah; do
  if [ $mode = auto ]; then unset RT_BVH; else export RT_BVH=$mode; fi
  timeout 600 python tools/kbench.py --stats --frames 20 --scenes c2,final,s14k,c4 >> gpurun_out/r2_kbench4.json 2>> gpurun_out/r2_kbench4.err
done
unset RT_BVH
RT_BUILD_TIMING=1 timeout 300 python tools/frame_once.py spheres_textured 500 1920 8 1 > gpurun_out/r2_buildtiming.log 2>&1
cat gpurun_out/r2_kbench4.json | cut -c1-330; cat gpurun_out/r2_buildtiming.log
