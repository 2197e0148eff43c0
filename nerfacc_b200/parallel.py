"""Ray-sharded data parallelism (one process per GPU).

nerfacc itself has no multi-GPU code.  Every ray is
independent on this path -- traversal, the segmented scans and the accumulation
never mix rays -- so rank k of P takes a contiguous ray shard, all packed outputs
stay rank-local (ray_indices are 0-based per shard), the occupancy grid is
replicated, and the only exchange is one all-reduce of the scalar loss.
"""
from typing import Tuple

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


class LossReduction:
    """Handle of a (possibly not yet started) all-reduce of the scalar loss; see :func:`all_reduce_loss_async`."""

    def __init__(self, value: torch.Tensor, scale: float, reduce: bool):
        self._value, self._scale = value, scale
        self._work = None
        self._started = not reduce

    def _start(self) -> None:
        if not self._started:
            self._started = True
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
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self._scale != 1.0:
                self._value *= self._scale
        return self._value


def all_reduce_loss_async(loss: torch.Tensor, average: bool = True, defer: bool = False) -> LossReduction:
    """Start the sum (or mean) of the per-rank scalar loss and return a handle.

    Call it as soon as the loss exists -- before ``backward()`` -- and read ``result()`` when the number is
    needed (logging, usually a step later): the 4-byte exchange then runs on NCCL's stream next to the backward
    kernels instead of stalling the compute stream until the slowest rank arrives.  With ``defer`` even the
    host-side enqueue (tens of microseconds of c10d / NCCL launch work, on a step that is host-bound) is parked
    until the next ``sampling()`` call waits for its march, where the host is idle anyway; ``result()`` starts
    it if no sampling call came first.  The value is a detached copy; the local autograd graph is untouched.
    """
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return LossReduction(loss.detach(), 1.0, reduce=False)
    handle = LossReduction(loss.detach().clone(), 1.0 / dist.get_world_size() if average else 1.0, reduce=True)
    if defer:
        _lib.idle_tasks.append(handle._start)
    else:
        handle._start()
    return handle


def all_reduce_loss(loss: torch.Tensor, average: bool = True) -> torch.Tensor:
    """Sum (or mean) of the per-rank scalar loss, complete on return (stream-ordered on CUDA)."""
    return all_reduce_loss_async(loss, average).result()
