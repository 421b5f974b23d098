set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests13.log 2>&1; tail -3 gpurun_out/r2_tests13.log
timeout 300 python tools/kbench.py --frames 40 --scenes c2,cornell,final,c4 > gpurun_out/r2_kbench13.json 2> gpurun_out/r2_kbench13.err
cut -c1-250 gpurun_out/r2_kbench13.json
