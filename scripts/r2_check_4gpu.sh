cd $GRAFT_REPO_ROOT
timeout 55 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
grep "local ms/step" gpurun_out/r2_bench_n4.err | tr '\n' ' '; cut -c1-200 gpurun_out/r2_bench_n4.json
