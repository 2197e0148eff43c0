# round-2 ncu evidence: launch list of the bench command + one --set full capture of every nfa kernel
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2}
NFA_BENCH_CLOCK_LOAD_STEPS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:march_kernel|offsets_kernel|expand_runs|composite_|vis_|generic_traverse|scan_|pack_|occ_|importance_sampling|accumulate_|intersect_sorted' -c 70 -f -o gpurun_out/${TAG}_kernels \
    python scripts/profile_kernels.py all > gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_prof.log
