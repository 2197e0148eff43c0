set -x
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k 'regex:march_kernel|offsets_kernel|expand_runs|composite_|vis_|generic_traverse|scan_|pack_|occ_|importance_sampling|accumulate_|intersect_sorted' -c 70 -f -o gpurun_out/r2_kernels python scripts/profile_kernels.py all > gpurun_out/r2_prof.log 2>&1
tail -n 3 gpurun_out/r2_prof.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json | cut -c1-700
