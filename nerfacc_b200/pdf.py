"""Importance sampling along rays and per-ray ``searchsorted``.

Public functions and argument meaning follow /root/reference/nerfacc/pdf.py:12-131; the native
work is ``nfa_importance_sampling`` / ``nfa_searchsorted`` (include/nerfacc_b200.h), one launch
each.  ``_sample_from_weighted`` is the reference's pure-torch cross-check (pdf.py:134-218), kept
because its tests and users compare against it.
"""
from typing import Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .data_specs import RayIntervals, RaySamples

Segments = Union[RayIntervals, RaySamples]


def _layout(seg: Segments, what: str):
    """(vals, packed_info | None, n_rays, edges_per_ray) of a batched or flattened operand."""
    vals = seg.vals
    _lib.require_cuda(vals, what)
    if vals.dtype != torch.float32:
        vals = vals.float()
    vals = vals.contiguous()
    if vals.dim() > 1:  # batched [..., E]
        e = vals.shape[-1]
        return vals, None, (vals.numel() // e if e else 0), e
    if seg.packed_info is None:
        # reference: RaySegmentsSpec::check() requires chunk_starts / chunk_cnts for flattened data
        raise RuntimeError(f"{what}: flattened data needs `packed_info`.")
    packed = seg.packed_info.to(torch.int64).contiguous()
    return vals, packed, packed.shape[0], 0


def _philox_state(device) -> Tuple[int, int]:
    """(seed, offset) of the device's default generator, advanced by 4 draws per thread.

    The reference takes this state on every call (``gen->philox_cuda_state(4)``, pdf.cu:309-317,
    :374-382), stratified or not, so the generator stream stays aligned with it call for call."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if len(torch.cuda.default_generators) <= idx:
        torch.cuda.init()
    gen = torch.cuda.default_generators[idx]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + 4)
    return seed & 0xFFFFFFFFFFFFFFFF, offset


def searchsorted(sorted_sequence: Segments, values: Segments) -> Tuple[Tensor, Tensor]:
    """``ids_left, ids_right`` with ``sorted_sequence.vals[ids_left] <= values.vals < sorted_sequence.vals[ids_right]``
    per ray; out-of-range values behave as if clipped to the ray's range (reference pdf.py:12-61)."""
    q_vals, q_packed, n_rays_q, q_edges = _layout(values, "searchsorted")
    k_vals, k_packed, n_rays_k, k_edges = _layout(sorted_sequence, "searchsorted")
    ids_left = torch.empty(q_vals.shape, dtype=torch.int64, device=q_vals.device)
    ids_right = torch.empty_like(ids_left)
    if q_vals.numel() == 0:
        return ids_left, ids_right
    q_ray = None
    if q_packed is not None and values.ray_indices is not None:
        q_ray = values.ray_indices.to(torch.int64).contiguous()
    _lib.call("nfa_searchsorted", q_vals.device, q_vals.numel(), q_vals.data_ptr(), _lib.ptr(q_packed),
              _lib.ptr(q_ray), n_rays_q, q_edges, k_vals.data_ptr(), _lib.ptr(k_packed), k_edges,
              ids_left.data_ptr(), ids_right.data_ptr())
    return ids_left, ids_right


def _importance_sampling(intervals: RayIntervals, cdfs: Tensor, n_intervals_per_ray: Union[Tensor, int],
                         stratified: bool, stot: Optional[Tuple[float, float, bool]] = None):
    vals, packed, n_rays, in_edges = _layout(intervals, "importance_sampling")
    cdfs = cdfs.contiguous()
    if cdfs.dtype != torch.float32:
        cdfs = cdfs.float()
    if cdfs.numel() != vals.numel():
        raise RuntimeError("importance_sampling: cdfs must have one value per edge of `intervals`.")
    dev = vals.device
    seed, offset = _philox_state(dev)
    max_in = 0
    if packed is not None:
        max_in = int(packed[:, 1].max()) if n_rays else 0  # sizes the per-ray staging area (one sync)

    if not isinstance(n_intervals_per_ray, Tensor):
        n = int(n_intervals_per_ray)
        lead = tuple(vals.shape[:-1]) if packed is None else (n_rays,)
        s_vals = torch.empty(lead + (n,), dtype=torch.float32, device=dev)
        e_vals = torch.empty(lead + (n + 1,), dtype=torch.float32, device=dev)
        t_starts = t_ends = None
        s_min = s_max = 0.0
        lindisp = False
        if stot is not None:
            s_min, s_max, lindisp = stot
            t_starts, t_ends = torch.empty_like(s_vals), torch.empty_like(s_vals)
        if n_rays and n:
            _lib.call("nfa_importance_sampling", dev, n_rays, vals.data_ptr(), cdfs.data_ptr(), _lib.ptr(packed),
                      in_edges, max_in, None, None, n, n, int(bool(stratified)), seed, offset, s_vals.data_ptr(), None,
                      e_vals.data_ptr(), None, None, None, _lib.ptr(t_starts), _lib.ptr(t_ends), s_min, s_max,
                      int(lindisp))
        return RayIntervals(vals=e_vals), RaySamples(vals=s_vals), t_starts, t_ends

    # per-ray counts -> flattened outputs (the layout the reference documents, pdf.py:88-104)
    cnts = n_intervals_per_ray.to(device=dev, dtype=torch.int64).contiguous().reshape(-1)
    if cnts.numel() != n_rays:
        raise RuntimeError("importance_sampling: n_intervals_per_ray must have one entry per ray.")
    e_cnts = (cnts + 1) * (cnts > 0)
    s_packed = torch.stack([torch.cumsum(cnts, 0) - cnts, cnts], -1).contiguous()
    e_packed = torch.stack([torch.cumsum(e_cnts, 0) - e_cnts, e_cnts], -1).contiguous()
    n_s, max_out = (int(cnts.sum()), int(cnts.max())) if n_rays else (0, 0)
    n_e = n_s + int((cnts > 0).sum()) if n_rays else 0
    s_vals = torch.empty(n_s, dtype=torch.float32, device=dev)
    s_ray = torch.empty(n_s, dtype=torch.int64, device=dev)
    e_vals = torch.empty(n_e, dtype=torch.float32, device=dev)
    e_ray = torch.empty(n_e, dtype=torch.int64, device=dev)
    e_left = torch.empty(n_e, dtype=torch.bool, device=dev)
    e_right = torch.empty(n_e, dtype=torch.bool, device=dev)
    if n_s:
        _lib.call("nfa_importance_sampling", dev, n_rays, vals.data_ptr(), cdfs.data_ptr(), _lib.ptr(packed), in_edges,
                  max_in, s_packed.data_ptr(), e_packed.data_ptr(), 0, max_out, int(bool(stratified)), seed, offset,
                  s_vals.data_ptr(), s_ray.data_ptr(), e_vals.data_ptr(), e_ray.data_ptr(), e_left.data_ptr(),
                  e_right.data_ptr(), None, None, 0.0, 0.0, 0)
    return (RayIntervals(vals=e_vals, packed_info=e_packed, ray_indices=e_ray, is_left=e_left, is_right=e_right),
            RaySamples(vals=s_vals, packed_info=s_packed, ray_indices=s_ray), None, None)


def importance_sampling(
    intervals: RayIntervals,
    cdfs: Tensor,
    n_intervals_per_ray: Union[Tensor, int],
    stratified: bool = False,
) -> Tuple[RayIntervals, RaySamples]:
    """Inverse-transform sampling of new intervals from per-edge CDFs (reference pdf.py:64-131).

    An ``int`` count gives batched outputs ``(n_rays, n + 1)`` / ``(n_rays, n)``; a per-ray count
    tensor gives flattened outputs with ``packed_info``, ``ray_indices``, ``is_left``, ``is_right``.
    With ``stratified`` every ray gets one jitter drawn from torch's CUDA generator, the same Philox
    stream position the reference uses.
    """
    out_iv, out_s, _, _ = _importance_sampling(intervals, cdfs, n_intervals_per_ray, stratified)
    return out_iv, out_s


def _sample_from_weighted(
    bins: Tensor,
    weights: Tensor,
    num_samples: int,
    stratified: bool = False,
    vmin: float = -torch.inf,
    vmax: float = torch.inf,
) -> Tuple[Tensor, Tensor]:
    """Pure-torch histogram resampling, bins (..., B+1), weights (..., B) -> edges (..., S+1), centres (..., S).

    Same construction as the reference's cross-check implementation (pdf.py:134-218, after mip-NeRF 360)."""
    n_bins = weights.shape[-1]
    assert bins.shape[-1] == n_bins + 1
    eps = torch.finfo(weights.dtype).eps
    pdf = torch.nn.functional.normalize(weights, p=1, dim=-1)
    zero, one = torch.zeros_like(pdf[..., :1]), torch.ones_like(pdf[..., :1])
    cdf = torch.cat([zero, torch.cumsum(pdf[..., :-1], dim=-1), one], dim=-1)

    opts = dict(dtype=bins.dtype, device=bins.device)
    if stratified:
        u_max = eps + (1 - eps) / num_samples
        max_jitter = (1 - u_max) / (num_samples - 1) - eps
        jitter = torch.rand(*bins.shape[:-1], 1, **opts) * max_jitter  # one jitter per ray
        u = torch.linspace(0, 1 - u_max, num_samples, **opts) + jitter
    else:
        pad = 1 / (2 * num_samples)
        u = torch.linspace(pad, 1 - pad - eps, num_samples, **opts).broadcast_to(bins.shape[:-1] + (num_samples,))

    hi = torch.searchsorted(cdf.contiguous(), u.contiguous(), side="right")
    lo = hi - 1
    cdf_lo, cdf_hi = cdf.gather(-1, lo), cdf.gather(-1, hi)
    bin_lo, bin_hi = bins.gather(-1, lo), bins.gather(-1, hi)
    frac = (u - cdf_lo) / torch.clamp(cdf_hi - cdf_lo, min=eps)
    centers = bin_lo + frac * (bin_hi - bin_lo)

    mids = (centers[..., 1:] + centers[..., :-1]) / 2
    first = (2 * centers[..., :1] - mids[..., :1]).clamp_min(vmin)
    last = (2 * centers[..., -1:] - mids[..., -1:]).clamp_max(vmax)
    return torch.cat([first, mids, last], dim=-1), centers
