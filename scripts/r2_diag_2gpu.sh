# N = 2 diagnostics: which part of the value arm costs the extra time
cd $GRAFT_REPO_ROOT
run() { # each run under its own timeout: a hung variant must not eat the call
   # name, env..., then bench args after --
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --no-cpu-baseline --no-reference-cuda "$@" > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
  echo "== $name"; grep "local ms/step" gpurun_out/r2s_$name.err | tr '\n' ' '; echo
}
run default X=1 -- --steps 50 --warmup 5
run steps20 X=1 -- --steps 20 --warmup 5
run nocollective NFA_BENCH_NO_COLLECTIVE=1 -- --steps 50 --warmup 5

run nccl NFA_BENCH_LOSS_TRANSPORT=nccl -- --steps 50 --warmup 5
run noload NFA_BENCH_CLOCK_LOAD_STEPS=0 -- --steps 50 --warmup 5
run nodefer NFA_BENCH_LOSS_DEFER=0 -- --steps 50 --warmup 5
