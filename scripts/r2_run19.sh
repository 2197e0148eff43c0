set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sampling or traverse or full_size" > gpurun_out/r2p_test.txt 2>&1
tail -3 gpurun_out/r2p_test.txt
for B in 8 16 32; do
  NFA_EXPAND_BATCH=$B timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size" > gpurun_out/r2p_test_b$B.txt 2>&1
  tail -1 gpurun_out/r2p_test_b$B.txt
  NFA_EXPAND_BATCH=$B python scripts/march_probe.py 20 > gpurun_out/r2p_probe_b$B.txt 2>&1
  cat gpurun_out/r2p_probe_b$B.txt
done
python scripts/march_probe.py 20 > gpurun_out/r2p_probe_auto.txt 2>&1
cat gpurun_out/r2p_probe_auto.txt
python scripts/march_trace.py > gpurun_out/r2p_trace.txt 2>&1
cat gpurun_out/r2p_trace.txt
