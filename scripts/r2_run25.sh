set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_gputest.txt 2>&1
tail -3 gpurun_out/r2u_gputest.txt
grep -q " passed" gpurun_out/r2u_gputest.txt && ! grep -q "failed\|error" gpurun_out/r2u_gputest.txt || exit 1
timeout 300 python scripts/march_probe.py 20 > gpurun_out/r2u_probe.txt 2>&1
cat gpurun_out/r2u_probe.txt
NFA_BENCH_CLOCK_LOAD_STEPS=300 timeout 600 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
cut -c1-300 gpurun_out/r2u_bench.json
