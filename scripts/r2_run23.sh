set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sampling or traverse or full_size" > gpurun_out/r2t_test.txt 2>&1
tail -3 gpurun_out/r2t_test.txt
timeout 300 python scripts/march_probe.py 20 > gpurun_out/r2t_probe.txt 2>&1
cat gpurun_out/r2t_probe.txt
