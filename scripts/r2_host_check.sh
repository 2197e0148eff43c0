# host-path micro-optimisations: full GPU tests, then the bench without the CPU legs (value / e2e / pipelined arms)
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2v_gputest.txt 2>&1
tail -2 gpurun_out/r2v_gputest.txt
NFA_BENCH_CLOCK_LOAD_STEPS=300 timeout 300 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipelined ms', d['pipelined']['ms_per_step'], d['roofline']['stages_us'], d['roofline']['sampling_call_us'])
PY
