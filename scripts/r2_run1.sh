set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json
