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


def shard_bounds(n_rays: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous shard; sizes differ by at most one ray."""
    base, rem = divmod(n_rays, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, rank: int, world_size: int):
    b, e = shard_bounds(rays_o.shape[0], rank, world_size)
    return rays_o[b:e], rays_d[b:e]


def all_reduce_loss(loss: torch.Tensor, average: bool = True) -> torch.Tensor:
    """Sum (or mean) of the per-rank scalar loss; a detached copy, the local graph is untouched.

    4-byte payload: the cost is launch latency, so it is issued on the compute stream.
    """
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return loss.detach()
    out = loss.detach().clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    if average:
        out /= dist.get_world_size()
    return out
