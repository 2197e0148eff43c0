set -x
cd $GRAFT_REPO_ROOT
python scripts/march_probe.py 20 > gpurun_out/r2c_probe.json 2>&1
NFA_MARCH_TILE=224 python scripts/march_probe.py 20 > gpurun_out/r2c_probe_t224.json 2>&1
NFA_MARCH_TILE=128 python scripts/march_probe.py 20 > gpurun_out/r2c_probe_t128.json 2>&1
NFA_MARCH_TILE=64 python scripts/march_probe.py 20 > gpurun_out/r2c_probe_t64.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 1 -f -o gpurun_out/r2c_march python scripts/profile_kernels.py step > gpurun_out/r2c_prof.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
head -1 gpurun_out/r2c_probe*.json; tail -8 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_bench.json
