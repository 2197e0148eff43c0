# compute-sanitizer over the tests that launch the reworked march / expand kernels (small sizes)
cd $GRAFT_REPO_ROOT
K='test_sampling_ball_scene or test_sampling_many_runs or test_sampling_nested or test_sampling_random_scenes or test_sampling_edge_cases or test_sampling_begin_end or test_traverse_grids_generic_modes or test_estimator_update_native'
t0=$(date +%s)
timeout 230 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/r2_san_race.txt 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r2_san_race.txt
t1=$(date +%s); echo "elapsed $((t1-t0))"
if [ $((t1-t0)) -lt 200 ]; then
  timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/r2_san_mem.txt 2>&1
  echo "memcheck rc=$?"; tail -6 gpurun_out/r2_san_mem.txt
fi
