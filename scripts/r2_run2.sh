set -x
cd $GRAFT_REPO_ROOT
python scripts/march_probe.py 20 > gpurun_out/r2b_probe.json 2>&1
NFA_MARCH_BRICK_STEPS=0 python scripts/march_probe.py 20 > gpurun_out/r2b_probe_nobrick.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 2 -f -o gpurun_out/r2b_march python scripts/profile_kernels.py step > gpurun_out/r2b_prof.log 2>&1
NFA_MARCH_BRICK_STEPS=0 ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 2 -f -o gpurun_out/r2b_march_nobrick python scripts/profile_kernels.py step >> gpurun_out/r2b_prof.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
python scripts/host_overhead.py > gpurun_out/r2b_host.log 2>&1
cat gpurun_out/r2b_probe.json gpurun_out/r2b_probe_nobrick.json; tail -5 gpurun_out/r2b_pytest.log; head -8 gpurun_out/r2b_host.log
