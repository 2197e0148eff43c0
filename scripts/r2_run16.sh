set -x
cd $GRAFT_REPO_ROOT
NFA_MARCH_SPLIT=0 python scripts/march_trace.py > gpurun_out/r2n_trace_split0.txt 2>&1
cat gpurun_out/r2n_trace_split0.txt
