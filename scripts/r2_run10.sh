set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/r2j_pytest.log
NFA_EXTRA_ONLY=scans,f3 python scripts/extra_configs.py ours > gpurun_out/r2j_extra_ours.log 2>&1
for v in s4 s16; do NFA_LIB=$PWD/gpurun_variants/lib_$v.so NFA_EXTRA_ONLY=scans python scripts/extra_configs.py ours 2>&1 | grep scans > gpurun_out/r2j_scans_$v.log; NFA_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scan" 2>&1 | tail -n 2 >> gpurun_out/r2j_scans_$v.log; done
tail -n 6 gpurun_out/r2j_pytest.log; grep -h "scans\|f3" gpurun_out/r2j_extra_ours.log gpurun_out/r2j_scans_*.log | cut -c1-420
