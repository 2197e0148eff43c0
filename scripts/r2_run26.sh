set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest.txt 2>&1
tail -2 gpurun_out/r2_gputest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 200 python scripts/march_probe.py 20 2>&1 | cut -c1-260
