set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_tests1.log
timeout 600 python tools/audit_configs.py c2 c3 c3n c4 c5 --samples gpurun_out/r2_audit_samples.npz > gpurun_out/r2_audit1.json 2> gpurun_out/r2_audit1.err
timeout 900 python tools/kbench.py --all --stats --frames 30 --scenes c2,cornell,final,c4 > gpurun_out/r2_kbench1.json 2> gpurun_out/r2_kbench1.err
RT_FUSED_GENERATE=0 timeout 300 python tools/kbench.py --frames 30 --scenes c2,cornell,final,c4 > gpurun_out/r2_kbench1_unfused.json 2>> gpurun_out/r2_kbench1.err
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shade|k_tail" -s 3 -c 3 -f -o gpurun_out/r2_prof_c4 python tools/frame_once.py spheres_textured 500 1920 8 3 > gpurun_out/r2_ncu_c4.log 2>&1
tail -3 gpurun_out/r2_tests1.log; cat gpurun_out/r2_audit1.json | cut -c1-600; cat gpurun_out/r2_kbench1.json
