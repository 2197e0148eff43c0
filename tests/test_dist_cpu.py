"""world_size-2 gloo check of the ray-sharding helpers (host logic of the N>1 path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerfacc_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays = torch.arange(n_rays * 3, dtype=torch.float32).reshape(n_rays, 3)
    o, d = parallel.shard_rays(rays, -rays, rank, world)
    local = o.sum()  # stands in for the per-rank loss
    tot = parallel.all_reduce_loss(local, average=False)
    handle = parallel.all_reduce_loss_async(local, average=True)  # started early, read late
    assert abs(float(handle.result()) * world - float(tot)) < 1e-3 * max(1.0, abs(float(tot)))
    assert float(handle.result()) == float(handle.result())  # idempotent
    # deferred enqueue: parked until the library is about to wait for the GPU (or until the value is asked for)
    from nerfacc_b200 import _lib
    h1 = parallel.all_reduce_loss_async(local, average=False, defer=True)
    h2 = parallel.all_reduce_loss_async(local * 2, average=False, defer=True)
    assert len(_lib.idle_tasks) == 2
    assert abs(float(h2.result()) - 2 * float(tot)) < 1e-3 * max(1.0, abs(float(tot)))  # runs h1 first, then h2
    assert not _lib.idle_tasks
    assert abs(float(h1.result()) - float(tot)) < 1e-3 * max(1.0, abs(float(tot)))
    h3 = parallel.all_reduce_loss_async(local, average=False, defer=True)
    _lib.run_idle_tasks()  # what sampling() does while the march runs
    assert abs(float(h3.result()) - float(tot)) < 1e-3 * max(1.0, abs(float(tot)))
    q.put((rank, o.shape[0], float(o[0, 0]) if o.shape[0] else -1.0, float(local), float(tot)))
    dist.destroy_process_group()


def test_shard_bounds_partition():
    for n in [0, 1, 7, 65536, 524288 + 3]:
        for w in [1, 2, 3, 8]:
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_allreduce():
    world, n_rays = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rays, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = float(torch.arange(n_rays * 3, dtype=torch.float32).sum())
    assert res[0][1] + res[1][1] == n_rays and res[0][2] == 0.0 and res[1][2] == 501 * 3.0
    assert abs(res[0][3] + res[1][3] - total) / total < 1e-6
    assert res[0][4] == res[1][4] and abs(res[0][4] - total) / total < 1e-6


def test_balanced_bounds_equalise_samples():
    # rank 0's rays are 3x as dense as rank 1's: the cut moves left until both halves hold the same samples
    cuts = parallel.balanced_bounds([0, 500, 1000], [3000, 1000])
    assert cuts == [0, 333, 1000]
    dens = [6.0] * 500 + [2.0] * 500
    assert abs(sum(dens[:cuts[1]]) - sum(dens[cuts[1]:])) <= 6.0
    # already balanced -> unchanged; all-empty -> even split; rescaled batch keeps the proportions
    assert parallel.balanced_bounds([0, 500, 1000], [100, 100]) == [0, 500, 1000]
    assert parallel.balanced_bounds([0, 10, 1000], [0, 0]) == [0, 500, 1000]
    assert parallel.balanced_bounds([0, 500, 1000], [3000, 1000], n_rays=2000) == [0, 667, 2000]
    c8 = parallel.balanced_bounds(list(range(0, 801, 100)), [1, 1, 1, 1, 5, 1, 1, 1])
    assert c8[0] == 0 and c8[-1] == 800 and all(a <= b for a, b in zip(c8, c8[1:]))
    w = [1] * 400 + [5] * 100 + [1] * 300
    per = [sum(w[a:b]) for a, b in zip(c8, c8[1:])]
    assert max(per) - min(per) <= 10  # within a couple of rays' worth of samples of each other


def _balancer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_rays = 1000
    density = torch.cat([torch.full((500,), 6.0), torch.full((500,), 2.0)])  # samples per ray
    bal = parallel.ShardBalancer(n_rays)
    history = []
    for _ in range(3):
        b, e = bal.bounds()
        local = int(density[b:e].sum())
        history.append(local)
        bal.update(local)
    # replicated occupancy grids: element-wise max over the ranks
    occs = torch.tensor([0.1, -1.0, 0.5, 0.0]) if rank == 0 else torch.tensor([0.3, -1.0, 0.2, 0.0])
    parallel.all_reduce_max_(occs)
    q.put((rank, history, bal.cuts, occs.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_balancer_and_grid_sync():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_balancer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, h0, cuts0, occ0), (_, h1, cuts1, occ1) = res
    assert cuts0 == cuts1                       # every rank derives the same boundaries
    assert h0[0] == 3000 and h1[0] == 1000      # split by rays: 3x imbalance
    assert abs(h0[-1] - h1[-1]) <= 8            # split by samples: balanced after one update
    assert occ0 == occ1 == [0.30000001192092896, -1.0, 0.5, 0.0]
