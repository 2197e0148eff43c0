# N=1 vs N=2 on one box (weak scaling of bench.py; also shows run-to-run spread)
cd $GRAFT_REPO_ROOT
one() { python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('N=1 value %.4f ms  e2e %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
two() { NFA_BENCH_LOSS_LAG=$1 NFA_BENCH_LOSS_TRANSPORT=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('N=2 lag=$1 $3 value %.4f ms  e2e %.4f ms  [%s]' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['loss_all_reduce']))"; }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29510 scripts/peer_check.py 2>&1 | grep -v "^\*\|OMP" | tail -4
one; two 1 29511 peer; two 2 29512 peer; two 1 29513 nccl; one; two 1 29514 peer; two 2 29515 peer
