set -x
mkdir -p gpurun_out
RT_EXTEND_MUX=0 timeout 300 python tools/kbench.py --frames 30 --scenes c2,cornell 2>> gpurun_out/r2_kbench11.err | sed "s/\"lib\": \"default\"/\"lib\": \"plain\"/" >> gpurun_out/r2_kbench11.json
timeout 600 python tools/mux_check.py > gpurun_out/r2_mux_check2.log 2>&1; tail -1 gpurun_out/r2_mux_check2.log
RT_EXTEND_MUX=1 timeout 900 python tools/kbench.py --all --frames 30 --scenes c2,cornell >> gpurun_out/r2_kbench11.json 2>> gpurun_out/r2_kbench11.err
cut -c1-240 gpurun_out/r2_kbench11.json
