set -x
cd $GRAFT_REPO_ROOT
export NFA_MARCH_SPLIT=0
for L in 4 8 16 32; do
  NFA_EXPAND_LANES=$L timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sampling or traverse or full_size" > gpurun_out/r2o_test_l$L.txt 2>&1
  tail -3 gpurun_out/r2o_test_l$L.txt
  NFA_EXPAND_LANES=$L python scripts/march_probe.py 20 > gpurun_out/r2o_probe_l$L.txt 2>&1
  cat gpurun_out/r2o_probe_l$L.txt
done
python scripts/march_probe.py 20 > gpurun_out/r2o_probe_auto.txt 2>&1
cat gpurun_out/r2o_probe_auto.txt
