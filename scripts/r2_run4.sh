set -x
cd $GRAFT_REPO_ROOT
python scripts/march_probe.py 20 > gpurun_out/r2d_probe_v1.json 2>&1
for v in v2 v3 v4; do
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python scripts/march_probe.py 20 > gpurun_out/r2d_probe_$v.json 2>&1
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sampling or traverse or full_size" 2>&1 | tail -3 > gpurun_out/r2d_pytest_$v.log
done
for k in 1 3 4; do NFA_MARCH_BRICK_WARPS=$k NFA_LIB=$PWD/gpurun_variants/lib_v3.so python scripts/march_probe.py 20 > gpurun_out/r2d_probe_v3_k$k.json 2>&1; done
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
head -1 gpurun_out/r2d_probe_*.json; tail -3 gpurun_out/r2d_pytest*.log
