"""torchrun --nproc-per-node N scripts/peer_check.py : PeerMailbox vs dist.all_reduce on N GPUs of one node."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfacc_b200 import parallel
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
mb = parallel.PeerMailbox.get(dev)
assert mb is not None, "mailbox unavailable"
bad = 0
pend = []
for k in range(200):
    v = torch.tensor(float(rank + 1) * 0.37 + k * 1e-3, device=dev)
    ref = v.clone(); dist.all_reduce(ref)
    h = parallel.all_reduce_loss_async(v, average=False, transport="peer")
    pend.append((h, ref))
    if len(pend) > 2:
        h0, r0 = pend.pop(0)
        got = h0.result()
        if abs(float(got) - float(r0)) > 1e-5 * abs(float(r0)):
            bad += 1
for h0, r0 in pend:
    if abs(float(h0.result()) - float(r0)) > 1e-5 * abs(float(r0)):
        bad += 1
mb.check()
# timing of post + collect vs nccl
def timeit(fn, n=300):
    torch.cuda.synchronize(); dist.barrier()
    import time
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6
x = torch.tensor(1.0, device=dev)
t_peer = timeit(lambda: parallel.all_reduce_loss_async(x, transport="peer").result())
t_nccl = timeit(lambda: parallel.all_reduce_loss_async(x, transport="nccl").result())
print(f"[rank {rank}] mismatches={bad}  peer {t_peer:.1f} us/op   nccl {t_nccl:.1f} us/op", flush=True)
parallel.PeerMailbox.shutdown()
dist.destroy_process_group()
