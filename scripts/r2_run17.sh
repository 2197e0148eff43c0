set -x
cd $GRAFT_REPO_ROOT
./scripts/exp/_bin/launch_overhead > gpurun_out/r2n_launch_overhead.txt 2>&1
cat gpurun_out/r2n_launch_overhead.txt
