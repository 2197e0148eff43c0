set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do
  NFA_EXTRA_ONLY=none timeout 300 python scripts/extra_configs.py 2>&1 | grep "f1" | cut -c1-260
  NFA_MARCH_SLOTS=8 NFA_EXTRA_ONLY=none timeout 300 python scripts/extra_configs.py 2>&1 | grep "f1" | cut -c1-260
done
timeout 200 python scripts/march_trace.py > gpurun_out/r2_trace_final.txt 2>&1
cat gpurun_out/r2_trace_final.txt
