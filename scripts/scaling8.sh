# one node, N = 8 (and 4): bench.py weak scaling + PeerMailbox check across all pairs
cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 scripts/peer_check.py 2>&1 | grep -v "^\*\|OMP" | tail -$N | cut -c1-120
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('N=%d value %.4f ms (%.2f G samples/s)  e2e %.4f ms  [%s]' % (d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['e2e']['ms_per_step'], d['config']['loss_all_reduce']))"
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('N=1 value %.4f ms (%.2f G samples/s)  e2e %.4f ms' % (d['ms_per_step'], d['value']/1e9, d['e2e']['ms_per_step']))"
