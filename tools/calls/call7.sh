set -x
mkdir -p gpurun_out
R="timeout 200 python tools/ref_gpu_frame.py"
( $R spheres 2 64 2 1; $R cornell 0 64 2 1; REF_GPU_NO_BVH=1 $R spheres 11 128 8 1; $R spheres 11 128 1 1; REF_GPU_STACK=131072 $R spheres 11 128 8 1; $R spheres 5 128 8 1 ) > gpurun_out/r2_refgpu3.json 2> gpurun_out/r2_refgpu3.err
cat gpurun_out/r2_refgpu3.json; tail -5 gpurun_out/r2_refgpu3.err
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests7.log 2>&1; tail -3 gpurun_out/r2_tests7.log
RT_BUILD_TIMING=1 timeout 300 python tools/frame_once.py spheres_textured 500 1920 8 1 > gpurun_out/r2_buildtiming4.log 2>&1
cat gpurun_out/r2_buildtiming4.log
