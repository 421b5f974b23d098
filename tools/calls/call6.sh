set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or builder or fused" > gpurun_out/r2_tests6.log 2>&1; tail -3 gpurun_out/r2_tests6.log
timeout 300 python tools/kbench.py --frames 50 --scenes c2,cornell > gpurun_out/r2_kbench6.json 2> gpurun_out/r2_kbench6.err
timeout 300 python tools/kbench.py --frames 50 --scenes c2,cornell --graph >> gpurun_out/r2_kbench6.json 2>> gpurun_out/r2_kbench6.err
cut -c1-300 gpurun_out/r2_kbench6.json
timeout 600 python tools/ref_gpu_frame.py spheres 11 1920 8 3 > gpurun_out/r2_refgpu2.json 2> gpurun_out/r2_refgpu2.err; echo "refgpu exit $?"
cat gpurun_out/r2_refgpu2.json; tail -3 gpurun_out/r2_refgpu2.err
RT_BUILD_TIMING=1 timeout 300 python tools/frame_once.py spheres_textured 500 1920 8 1 > gpurun_out/r2_buildtiming3.log 2>&1
cat gpurun_out/r2_buildtiming3.log
cd real-time-ray-tracing-engine_b200/host && for i in 1 2; do ./raytracer --camera dynamic --headless --scene spheres --width 1920 --samples 1 --depth 8 --frames 300 2>&1 | tail -2; done
