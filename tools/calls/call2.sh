set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_tests2.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_tests2.log
timeout 600 python tools/audit_configs.py c2 c3 c3n c4 c5 --samples gpurun_out/r2_audit_samples2.npz > gpurun_out/r2_audit2.json 2> gpurun_out/r2_audit2.err
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench exit $?"
grep -E "passed|failed|error" gpurun_out/r2_tests2.log | tail -5; cut -c1-400 gpurun_out/r2_audit2.json; tail -c 3000 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
