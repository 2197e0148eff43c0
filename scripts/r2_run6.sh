set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/r2f_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python scripts/extra_configs.py ours > gpurun_out/r2f_extra_ours.log 2>&1
python scripts/extra_configs.py reference-cuda > gpurun_out/r2f_extra_ref.log 2>&1
bash scripts/r2_profile.sh r2f
tail -n 4 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_extra_ours.log | tail -n 12; cat gpurun_out/r2f_extra_ref.log | tail -n 12
