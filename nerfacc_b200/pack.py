"""pack_info: per-sample ray indices -> per-ray (start, count).

Mirrors /root/reference/nerfacc/pack.py:10-49 (same name, arguments, dtype rule
and error behaviour); the histogram + scan run in ``nfa_pack_info`` instead of
``index_add_`` + ``cumsum``.
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def _stash_packed_info(ray_indices: Tensor, packed_info: Tensor, n_rays: int) -> None:
    """Remember the packed_info a traversal produced together with `ray_indices`.

    Both come out of the same kernel, so consumers that are handed exactly this
    tensor object (unmodified) need not rebuild the segments from 8 B/sample of indices.
    """
    ray_indices._nfa_packed = (packed_info, int(n_rays), ray_indices._version)


def _stashed_packed_info(ray_indices: Tensor, n_rays: Optional[int]) -> Optional[Tensor]:
    st = getattr(ray_indices, "_nfa_packed", None)
    if st is None:
        return None
    packed_info, n, version = st
    if version != ray_indices._version or (n_rays is not None and int(n_rays) != n):
        return None
    return packed_info


@torch.no_grad()
def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    """(n_rays, 2) tensor of (chunk_start, chunk_cnt) for grouped `ray_indices`.

    Same contract as the reference: CUDA only (``NotImplementedError`` otherwise),
    output dtype follows the input dtype, ``n_rays=None`` infers ``max + 1``.
    """
    assert ray_indices.dim() == 1, "ray_indices must be a 1D tensor with shape (n_samples)."
    if not ray_indices.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    # (with n_rays=None the reference returns max + 1 rows, which the stash -- one row per ray -- need not have)
    cached = _stashed_packed_info(ray_indices, n_rays) if n_rays is not None else None
    if cached is not None:
        return cached if cached.dtype == ray_indices.dtype else cached.to(ray_indices.dtype)
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() > 0 else 0
    n_rays = int(n_rays)
    device = ray_indices.device
    idx = ray_indices.contiguous()
    if idx.dtype != torch.int64:
        idx = idx.to(torch.int64)
    lib = _lib.load()
    out = torch.empty((n_rays, 2), dtype=torch.int64, device=device)
    if n_rays > 0:
        ws = torch.empty(lib.nfa_pack_info_workspace_bytes(n_rays), dtype=torch.uint8, device=device)
        _lib.call("nfa_pack_info", device, idx.numel(), _lib.ptr(idx), n_rays, _lib.ptr(out), _lib.ptr(ws))
    return out if ray_indices.dtype == torch.int64 else out.to(ray_indices.dtype)
