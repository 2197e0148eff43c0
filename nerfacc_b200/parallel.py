"""Ray-sharded data parallelism (one process per GPU).

nerfacc itself has no multi-GPU code.  Every ray is
independent on this path -- traversal, the segmented scans and the accumulation
never mix rays -- so rank k of P takes a contiguous ray shard, all packed outputs
stay rank-local (ray_indices are 0-based per shard), the occupancy grid is
replicated, and the only exchange is one all-reduce of the scalar loss.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_rays: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous shard; sizes differ by at most one ray."""
    base, rem = divmod(n_rays, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, rank: int, world_size: int):
    b, e = shard_bounds(rays_o.shape[0], rank, world_size)
    return rays_o[b:e], rays_d[b:e]


def balanced_bounds(prev_bounds, samples_per_rank, n_rays: Optional[int] = None):
    """Re-split a ray batch so that every rank gets the same number of SAMPLES, not rays.

    `prev_bounds` (P + 1 increasing ints) are the shard boundaries of the last step and `samples_per_rank[k]` the
    samples rank k emitted for its shard.  Rays of one shard are taken to be equally dense (samples per ray), which
    makes the cumulative sample count piecewise linear in the ray index; the new boundaries are where it crosses
    k / P of the total.  `n_rays` rescales to a batch of another size (the reference's training loop resizes its
    batch every step, examples/train_ngp_nerf_occ.py:150-203).  Pure host arithmetic, identical on every rank."""
    P = len(samples_per_rank)
    assert len(prev_bounds) == P + 1
    lo, hi = float(prev_bounds[0]), float(prev_bounds[-1])
    total = float(sum(samples_per_rank))
    n_old = hi - lo
    n_new = int(n_rays) if n_rays is not None else int(round(n_old))
    if total <= 0 or n_old <= 0:
        return [shard_bounds(n_new, k, P)[0] for k in range(P)] + [n_new]
    cum = [0.0]
    for s_k in samples_per_rank:
        cum.append(cum[-1] + float(s_k))
    out, seg = [0], 0
    for k in range(1, P):
        target = total * k / P
        while seg < P - 1 and cum[seg + 1] < target:
            seg += 1
        width = float(prev_bounds[seg + 1] - prev_bounds[seg])
        dens = (cum[seg + 1] - cum[seg]) / width if width > 0 else 0.0
        x = float(prev_bounds[seg]) + ((target - cum[seg]) / dens if dens > 0 else width)
        cut = int(round((x - lo) / n_old * n_new))
        out.append(min(max(cut, out[-1]), n_new))
    out.append(n_new)
    return out


class ShardBalancer:
    """Keeps the ray shards of a data-parallel job balanced by samples: call ``update(n_samples)`` once per step
    with the local sample count (one tiny all-gather) and take ``bounds(rank)`` for the next batch."""

    def __init__(self, n_rays: int, world_size: Optional[int] = None, rank: Optional[int] = None):
        ready = dist.is_available() and dist.is_initialized()
        self.world = world_size if world_size is not None else (dist.get_world_size() if ready else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if ready else 0)
        self.cuts = [shard_bounds(n_rays, k, self.world)[0] for k in range(self.world)] + [n_rays]

    def bounds(self, rank: Optional[int] = None) -> Tuple[int, int]:
        k = self.rank if rank is None else rank
        return self.cuts[k], self.cuts[k + 1]

    def update(self, n_samples_local: int, n_rays: Optional[int] = None, device=None):
        counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        counts[self.rank] = int(n_samples_local)
        if self.world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        self.cuts = balanced_bounds(self.cuts, counts.tolist(), n_rays)
        return self.cuts


def all_reduce_max_(t: torch.Tensor) -> torch.Tensor:
    """In-place element-wise maximum over the ranks (no-op on one rank): keeps the replicated occupancy grids of
    a ray-sharded job identical after `OccGridEstimator._update` (set `estimator.sync_across_ranks`)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


class PeerMailbox:
    """Sum of per-rank scalars over NVLink peer memory (csrc/peer.cu), for the ranks of ONE node.

    Every rank owns a mailbox in its device memory, exported with CUDA IPC and mapped by its peers; ``post`` stores
    this rank's value into every mailbox with one tiny kernel, ``collect`` sums a turn of the local mailbox with
    another.  No NCCL / c10d call per step: on the host-bound step of this path the c10d route cost ~75 us per
    step at N = 2, the two launches here ~10 us.  Built once (``PeerMailbox.get(device)``, collective: every rank
    must call it), it falls back to ``None`` when the handles cannot be exchanged or mapped.
    """

    TURNS = 16  # ring of turns: a rank is never more than the consumer's lag (a few steps) ahead of the slowest
    _instances: dict = {}

    def __init__(self, device: torch.device):
        """Local part only (allocate + export this rank's mailbox); `get` drives the collective phases."""
        import ctypes as C
        self.device, self.world, self.rank = device, dist.get_world_size(), dist.get_rank()
        self._C, self.step = C, 0
        self.peers, self.table, self.status = [], None, None
        lib = _lib.load()
        box, handle = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            _lib.check(lib.nfa_mailbox_create(self.world, self.TURNS, C.byref(box), handle), "nfa_mailbox_create")
        self.box, self.handle = box.value, bytes(handle)

    def _open_peers(self, handles) -> None:
        """Local part: map every other rank's mailbox."""
        C, lib, ptrs = self._C, _lib.load(), []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.box)
                continue
            peer = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.nfa_mailbox_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(peer)),
                           "nfa_mailbox_open")
            self.peers.append(peer.value)
            ptrs.append(peer.value)
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _release(self) -> None:
        lib = _lib.load()
        for peer in self.peers:
            lib.nfa_mailbox_close(self._C.c_void_p(peer))
        self.peers = []
        if self.box:
            lib.nfa_mailbox_destroy(self._C.c_void_p(self.box))
            self.box = None

    @classmethod
    def get(cls, device: torch.device):
        """The mailbox of this process for `device` (built on first use; collective).  None if unavailable.

        Set-up alternates local steps (which may fail on one rank only: IPC not permitted, a rank on another
        node, ...) with collectives; after every local step the ranks agree on an all-reduced ok flag BEFORE
        anyone enters the next collective, so a one-sided failure ends in the same clean fallback (None -> NCCL)
        on every rank instead of mismatched collectives."""
        key = (device.type, device.index)
        if key in cls._instances:
            return cls._instances[key]

        def agreed(ok: bool) -> bool:
            flag = torch.full((1,), 1.0 if ok else 0.0, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return bool(flag.item())

        inst = None
        try:  # phase 1 (local): allocate and export
            inst = cls(device)
        except Exception:
            inst = None
        if not agreed(inst is not None):
            if inst is not None:
                inst._release()
            cls._instances[key] = None
            return None
        handles = [None] * inst.world
        dist.all_gather_object(handles, inst.handle)  # phase 2 (collective): every rank is known to be in it
        try:  # phase 3 (local): map the peers
            inst._open_peers(handles)
            mapped = True
        except Exception:
            mapped = False
        if not agreed(mapped):
            torch.cuda.synchronize(device)
            dist.barrier()  # nobody frees a mailbox a peer may still be mapping
            inst._release()
            cls._instances[key] = None
            return None
        dist.barrier()  # nobody posts before every mailbox is mapped everywhere
        cls._instances[key] = inst
        return inst

    def check(self) -> None:
        """Raise if a ``collect`` gave up waiting for a rank (synchronises the device)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("PeerMailbox: a rank's value did not arrive (a process died or fell far behind)")

    @classmethod
    def shutdown(cls) -> None:
        """Unmap the peers' mailboxes, then free the own one (collective; call before destroy_process_group)."""
        for inst in cls._instances.values():
            if inst is None:
                continue
            torch.cuda.synchronize(inst.device)
            lib = _lib.load()
            dist.barrier()
            for peer in inst.peers:
                lib.nfa_mailbox_close(inst._C.c_void_p(peer))
            dist.barrier()
            lib.nfa_mailbox_destroy(inst._C.c_void_p(inst.box))
        cls._instances.clear()

    def post(self, value: torch.Tensor):
        """Send `value` (a float32 scalar on this device) to every rank.  Returns the ticket for ``collect``."""
        self.step += 1
        turn, tag = self.step % self.TURNS, self.step & 0xFFFFFFFF
        _lib.call("nfa_mailbox_post", self.device, value.data_ptr(), self.table.data_ptr(), self.world, self.rank, turn,
                  tag)
        return turn, tag

    def collect(self, ticket, scale: float = 1.0) -> torch.Tensor:
        out = torch.empty((), dtype=torch.float32, device=self.device)
        _lib.call("nfa_mailbox_sum", self.device, self.box, self.world, ticket[0], ticket[1], float(scale),
                  out.data_ptr(), self.status.data_ptr())
        return out


class LossReduction:
    """Handle of a (possibly not yet started) all-reduce of the scalar loss; see :func:`all_reduce_loss_async`."""

    def __init__(self, value: torch.Tensor, scale: float, reduce: bool, mailbox: Optional[PeerMailbox] = None):
        self._value, self._scale = value, scale
        self._work = None
        self._started = not reduce
        self._mailbox, self._ticket = mailbox, None

    def _start(self) -> None:
        if not self._started:
            self._started = True
            if self._mailbox is not None:
                self._ticket = self._mailbox.post(self._value)
            else:
                self._work = dist.all_reduce(self._value, op=dist.ReduceOp.SUM, async_op=True)

    def result(self) -> torch.Tensor:
        """The reduced loss.  Orders the current stream after the collective (no host sync on CUDA)."""
        if not self._started:
            if self._start in _lib.idle_tasks:
                # keep the collective order identical on every rank: run everything parked before this one
                while _lib.idle_tasks:
                    task = _lib.idle_tasks.pop(0)
                    task()
                    if task == self._start:
                        break
            else:
                self._start()
        if self._ticket is not None:
            self._value = self._mailbox.collect(self._ticket, self._scale)
            self._ticket = None
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self._scale != 1.0:
                self._value *= self._scale
        return self._value


def all_reduce_loss_async(loss: torch.Tensor, average: bool = True, defer: bool = False,
                          transport: str = "nccl") -> LossReduction:
    """Start the sum (or mean) of the per-rank scalar loss and return a handle.

    Call it as soon as the loss exists -- before ``backward()`` -- and read ``result()`` when the number is
    needed (logging, usually a step later): the 4-byte exchange then runs next to the backward kernels instead of
    stalling the compute stream until the slowest rank arrives.  With ``defer`` even the host-side enqueue is
    parked until the next ``sampling()`` call waits for its march, where the host is idle anyway; ``result()``
    starts it if no sampling call came first.  ``transport="peer"`` sends the scalar through :class:`PeerMailbox`
    (NVLink peer stores, single node, CUDA float32 scalar) instead of ``dist.all_reduce``; it falls back to NCCL
    when the mailbox cannot be set up.  The value is a detached copy; the local autograd graph is untouched.
    """
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return LossReduction(loss.detach(), 1.0, reduce=False)
    scale = 1.0 / dist.get_world_size() if average else 1.0
    mailbox = None
    if transport == "peer" and loss.is_cuda and loss.dtype == torch.float32 and loss.numel() == 1:
        mailbox = PeerMailbox.get(loss.device)
    elif transport not in ("nccl", "peer"):
        raise ValueError(f"unknown transport: {transport}")
    # the mailbox kernel reads the value when it runs: hand it the loss itself (kept alive by the handle)
    value = loss.detach() if mailbox is not None else loss.detach().clone()
    handle = LossReduction(value, scale, reduce=True, mailbox=mailbox)
    if defer:
        _lib.idle_tasks.append(handle._start)
    else:
        handle._start()
    return handle


def all_reduce_loss(loss: torch.Tensor, average: bool = True) -> torch.Tensor:
    """Sum (or mean) of the per-rank scalar loss, complete on return (stream-ordered on CUDA)."""
    return all_reduce_loss_async(loss, average).result()
