set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/r2m_pytest.log
for k in 64 32 0; do NFA_MARCH_SPLIT=$k python scripts/march_probe.py 20 > gpurun_out/r2m_probe_split$k.json 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 1 -f -o gpurun_out/r2m_march_split python scripts/profile_kernels.py step > gpurun_out/r2m_prof.log 2>&1
tail -n 4 gpurun_out/r2m_pytest.log; head -n 1 gpurun_out/r2m_probe_split*.json | cut -c1-200
