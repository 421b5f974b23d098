set -x
mkdir -p gpurun_out
for b in 0 8192 14336; do RT_EXTEND_DYN_SMEM=$b timeout 300 python tools/kbench.py --frames 30 --scenes c2 2>> gpurun_out/r2_kbench10.err | sed "s/\"lib\": \"default\"/\"lib\": \"dyn$b\"/" >> gpurun_out/r2_kbench10.json; done
cut -c1-250 gpurun_out/r2_kbench10.json
RT_EXTEND_MUX=1 RT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_extend_mux" -s 3 -c 3 -f -o gpurun_out/r2_prof_mux python tools/frame_once.py spheres 11 1920 8 2 > gpurun_out/r2_ncu_mux.log 2>&1; tail -3 gpurun_out/r2_ncu_mux.log
