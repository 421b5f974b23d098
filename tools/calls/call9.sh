set -x
mkdir -p gpurun_out
timeout 600 python tools/mux_check.py > gpurun_out/r2_mux_check.log 2>&1; cat gpurun_out/r2_mux_check.log
for m in 0 1; do RT_EXTEND_MUX=$m timeout 300 python tools/kbench.py --frames 40 --scenes c2,cornell,final,c4 2>> gpurun_out/r2_kbench9.err | sed "s/\"lib\": \"default\"/\"lib\": \"mux$m\"/" >> gpurun_out/r2_kbench9.json; done
cut -c1-250 gpurun_out/r2_kbench9.json
