# 2-GPU check of the N > 1 path (mailbox set-up in phases, LOSS_LAG 2, per-rank timings)
set -x
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref_n2.json 2>> gpurun_out/r2h_bench_n2.err
python -m pytest tests -m gpu -q -k "not_current or two_streams" 2>&1 | tail -n 5 > gpurun_out/r2h_pytest_2gpu.log
cat gpurun_out/r2h_bench_n2.json | cut -c1-1500; tail -n 5 gpurun_out/r2h_bench_n2.err; cat gpurun_out/r2h_pytest_2gpu.log
