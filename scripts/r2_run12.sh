set -x
cd $GRAFT_REPO_ROOT
timeout 900 bash scripts/r2_profile.sh r2
python scripts/host_overhead.py 2>&1 | head -n 6 > gpurun_out/r2_host_overhead.txt
python scripts/march_probe.py 20 > gpurun_out/r2_march_probe.json 2>&1
python scripts/composite_probe.py 20 > gpurun_out/r2_composite_probe.json 2>&1
ls -la gpurun_out | grep "r2_"
