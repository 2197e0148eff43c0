# final evidence of the round: full GPU tests, both bench arms, ncu launch list, ncu --set full of the step's kernels
set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest.txt 2>&1
tail -3 gpurun_out/r2_gputest.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
cut -c1-600 gpurun_out/r2_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cut -c1-800 gpurun_out/r2_bench.json
NFA_BENCH_CLOCK_LOAD_STEPS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:march_kernel|offsets_kernel|expand_runs|composite_' -c 10 -f -o gpurun_out/r2_step_kernels \
    python scripts/profile_kernels.py step > gpurun_out/r2_step_prof.log 2>&1
tail -3 gpurun_out/r2_step_prof.log
NFA_EXTRA_ONLY=c3 timeout 300 python scripts/extra_configs.py > gpurun_out/r2_extra_c3.txt 2>&1 || true
tail -5 gpurun_out/r2_extra_c3.txt
