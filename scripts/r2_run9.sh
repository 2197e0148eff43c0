set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scan or distortion or golden or rendering" 2>&1 | tail -n 8 > gpurun_out/r2i_pytest_scan.log
NFA_EXTRA_ONLY=scans python scripts/extra_configs.py ours > gpurun_out/r2i_extra_ours.log 2>&1
tail -n 8 gpurun_out/r2i_pytest_scan.log; tail -n 2 gpurun_out/r2i_extra_ours.log
