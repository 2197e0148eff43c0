set -x
cd $GRAFT_REPO_ROOT
python scripts/march_probe.py 20 > gpurun_out/r2l_probe_base.json 2>&1
for v in m4 m5; do
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python scripts/march_probe.py 20 > gpurun_out/r2l_probe_$v.json 2>&1
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sampling or traverse or full_size" 2>&1 | tail -n 2 > gpurun_out/r2l_pytest_$v.log
done
head -n 1 gpurun_out/r2l_probe_*.json | cut -c1-200; tail -n 1 gpurun_out/r2l_pytest_*.log
