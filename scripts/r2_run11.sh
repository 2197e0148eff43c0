set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r2k_pytest.log
NFA_EXTRA_ONLY=c4,f3 python scripts/extra_configs.py ours > gpurun_out/r2k_extra_ours.log 2>&1
tail -n 6 gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_extra_ours.log | cut -c1-460
