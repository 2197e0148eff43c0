set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sampling or traverse or full_size" > gpurun_out/r2q_test.txt 2>&1
tail -3 gpurun_out/r2q_test.txt
python scripts/march_probe.py 20 > gpurun_out/r2q_probe.txt 2>&1
cat gpurun_out/r2q_probe.txt
python scripts/march_trace.py > gpurun_out/r2q_trace.txt 2>&1
cat gpurun_out/r2q_trace.txt
NFA_BENCH_CLOCK_LOAD_STEPS=200 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -c 3000 gpurun_out/r2q_bench.json
