"""Differentiable volume rendering over packed or batched samples.

Public surface and semantics mirror /root/reference/nerfacc/volrend.py
(rendering :15-164, render_transmittance_from_* :167-278, render_weight_from_*
:281-376, render_visibility_from_* :379-494, accumulate_along_rays[_] :497-587).

Packed (flattened) CUDA inputs run through two fused kernels
(``nfa_composite_fwd`` / ``nfa_composite_bwd``); batched ``(n_rays, n_samples)``
inputs keep the reference's plain-torch formulation (cumsum / cumprod on the last
dim), which is also the only mode that works on CPU tensors -- exactly as in the
reference.
"""
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .pack import _stashed_packed_info, pack_info
from .scan import exclusive_prod, exclusive_sum


# --------------------------------------------------------------------------
# native autograd functions
# --------------------------------------------------------------------------

def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_contiguous():
        t = t.contiguous()
    if t.dtype != torch.float32:
        raise RuntimeError("nerfacc_b200 rendering kernels support float32 inputs only.")
    return t


def _segments(packed_info: Optional[Tensor], ray_indices: Optional[Tensor], n_rays: Optional[int]) -> Tensor:
    """(n_rays, 2) int64 contiguous segments from whichever addressing the caller gave."""
    if packed_info is None:
        packed_info = _stashed_packed_info(ray_indices, n_rays)
        if packed_info is None:
            packed_info = pack_info(ray_indices, n_rays)
    pi = packed_info.contiguous()
    if pi.dtype != torch.int64:
        pi = pi.to(torch.int64)
    return pi


def _composite_forward(dens, rgbs, packed_info, t_starts, t_ends, prefix_trans, bkgd, from_alpha: bool,
                       expected_depths: bool, want_rays: bool):
    """One launch of nfa_composite_fwd.  Returns (weights, trans, alphas, colors, opac, depths, raw)."""
    device = dens.device
    n = dens.shape[0]
    n_rays = packed_info.shape[0]
    weights = torch.empty_like(dens)
    trans = torch.empty_like(dens)
    alphas = dens if from_alpha else torch.empty_like(dens)
    if want_rays:
        colors = torch.empty((n_rays, 3), dtype=torch.float32, device=device) if rgbs is not None else None
        opac = torch.empty((n_rays, 1), dtype=torch.float32, device=device)
        depths = torch.empty((n_rays, 1), dtype=torch.float32, device=device) if t_starts is not None else None
        raw = torch.empty((n_rays, 5), dtype=torch.float32, device=device)
    else:
        colors = opac = depths = raw = None
    _lib.call("nfa_composite_fwd", device, n_rays, n, _lib.ptr(packed_info), _lib.ptr(t_starts), _lib.ptr(t_ends),
              _lib.ptr(dens), int(from_alpha), _lib.ptr(rgbs), _lib.ptr(prefix_trans), _lib.ptr(bkgd),
              int(expected_depths), _lib.ptr(weights), _lib.ptr(trans),
              None if from_alpha else _lib.ptr(alphas), _lib.ptr(colors), _lib.ptr(opac), _lib.ptr(depths),
              _lib.ptr(raw))
    if from_alpha:
        alphas = dens.detach()  # an input, handed back for symmetry; carries no graph edge
    return weights, trans, alphas, colors, opac, depths, raw


class _Composite(torch.autograd.Function):
    """Fused weights (+ optional per-ray accumulation) with a recompute backward."""

    @staticmethod
    def forward(ctx, dens, rgbs, packed_info, t_starts, t_ends, prefix_trans, bkgd,
                from_alpha: bool, expected_depths: bool, want_rays: bool):
        weights, trans, alphas, colors, opac, depths, raw = _composite_forward(
            dens, rgbs, packed_info, t_starts, t_ends, prefix_trans, bkgd, from_alpha, expected_depths, want_rays)
        ctx.from_alpha, ctx.expected_depths = from_alpha, expected_depths
        ctx.save_for_backward(dens, rgbs, packed_info, t_starts, t_ends, prefix_trans, bkgd, raw)
        ctx.set_materialize_grads(False)
        return weights, trans, alphas, colors, opac, depths

    @staticmethod
    def backward(ctx, g_w, g_t, g_a, g_c, g_o, g_d):
        dens, rgbs, packed_info, t_starts, t_ends, prefix_trans, bkgd, raw = ctx.saved_tensors
        device = dens.device
        need_dens, need_rgbs = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and rgbs is not None
        if not (need_dens or need_rgbs):
            return (None,) * 10
        g_dens = torch.empty_like(dens)
        g_rgbs = torch.empty_like(rgbs) if need_rgbs else None
        if ctx.from_alpha:
            g_a = None  # `alphas` output is the detached input
        _lib.call("nfa_composite_bwd", device, packed_info.shape[0], dens.shape[0], _lib.ptr(packed_info), _lib.ptr(t_starts),
                  _lib.ptr(t_ends), _lib.ptr(dens), int(ctx.from_alpha), _lib.ptr(rgbs), _lib.ptr(prefix_trans),
                  _lib.ptr(bkgd), int(ctx.expected_depths), _lib.ptr(raw),
                  _lib.ptr(_f32c(g_c)), _lib.ptr(_f32c(g_o)), _lib.ptr(_f32c(g_d)),
                  _lib.ptr(_f32c(g_w)), _lib.ptr(_f32c(g_t)), _lib.ptr(_f32c(g_a)),
                  _lib.ptr(g_dens), _lib.ptr(g_rgbs))
        return g_dens if need_dens else None, g_rgbs, None, None, None, None, None, None, None, None


class _Accumulate(torch.autograd.Function):
    """accumulate_along_rays for flattened samples (reference volrend.py:546-558)."""

    @staticmethod
    def forward(ctx, weights, values, ray_indices, packed_info, n_rays: int):
        device = weights.device
        dim = 1 if values is None else values.shape[-1]
        if packed_info is not None:
            out = torch.empty((n_rays, dim), dtype=torch.float32, device=device)
            _lib.call("nfa_accumulate_fwd", device, n_rays, _lib.ptr(packed_info), _lib.ptr(weights),
                      _lib.ptr(values), dim, _lib.ptr(out))
        else:
            out = torch.zeros((n_rays, dim), dtype=torch.float32, device=device)
            _lib.call("nfa_accumulate_atomic", device, weights.shape[0], _lib.ptr(ray_indices), _lib.ptr(weights),
                      _lib.ptr(values), dim, _lib.ptr(out))
        ctx.save_for_backward(weights, values, ray_indices)
        return out

    @staticmethod
    def backward(ctx, g_out):
        weights, values, ray_indices = ctx.saved_tensors
        device = weights.device
        dim = 1 if values is None else values.shape[-1]
        g_out = _f32c(g_out)
        g_w = torch.empty_like(weights) if ctx.needs_input_grad[0] else None
        g_v = torch.empty_like(values) if (values is not None and ctx.needs_input_grad[1]) else None
        if weights.shape[0] > 0 and (g_w is not None or g_v is not None):
            _lib.call("nfa_accumulate_bwd", device, weights.shape[0], _lib.ptr(ray_indices), _lib.ptr(weights),
                      _lib.ptr(values), dim, _lib.ptr(g_out), _lib.ptr(g_w), _lib.ptr(g_v))
        return g_w, g_v, None, None, None


def _packed_mode(x: Tensor, packed_info: Optional[Tensor], ray_indices: Optional[Tensor]) -> bool:
    return packed_info is not None or ray_indices is not None


def _grad_through(*tensors: Optional[Tensor]) -> bool:
    """Does the caller differentiate through any of these?  The fused kernels treat t_starts / t_ends /
    prefix_trans as constants (their backward returns None for them); the reference's op sequence
    (volrend.py:271-277) is differentiable in them, so such calls take the ATen formulation over the native
    scans instead of silently dropping the gradient."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


_uniform_segments = {}


def _batched_segments(x: Tensor, prefix_trans: Optional[Tensor]) -> Optional[Tensor]:
    """Batched (n_rays, S) CUDA input -> the (n_rays, 2) segments that address it as a packed array, so the
    batched flavour (the proposal estimator's) runs the same fused kernels.  None = keep the ATen formulation."""
    if not x.is_cuda or x.dim() != 2 or x.numel() == 0 or x.dtype != torch.float32 or prefix_trans is not None:
        return None
    key = (x.shape[0], x.shape[1], x.device)
    seg = _uniform_segments.get(key)
    if seg is None:
        if len(_uniform_segments) > 16:
            _uniform_segments.clear()
        starts = torch.arange(x.shape[0], device=x.device, dtype=torch.int64) * x.shape[1]
        seg = torch.stack([starts, torch.full_like(starts, x.shape[1])], -1).contiguous()
        _uniform_segments[key] = seg
    return seg


def _composite_packed(dens: Tensor, rgbs: Optional[Tensor], t_starts: Optional[Tensor], t_ends: Optional[Tensor],
                      packed_info: Optional[Tensor], ray_indices: Optional[Tensor], n_rays: Optional[int],
                      prefix_trans: Optional[Tensor], bkgd: Optional[Tensor], from_alpha: bool,
                      expected_depths: bool, want_rays: bool):
    _lib.require_cuda(dens, "rendering")
    assert dens.dim() == 1, "flattened inputs must be 1-D"
    pi = _segments(packed_info, ray_indices, n_rays)
    if dens.numel() == 0:
        # nothing to launch; per-ray outputs of empty rays are zero (+ background)
        if not want_rays:
            return dens, dens, dens, None, None, None
        zeros = torch.zeros((pi.shape[0], 1), dtype=torch.float32, device=dens.device)
        colors = zeros.expand(-1, 3).clone() if rgbs is not None else None
        if colors is not None and bkgd is not None:
            colors = colors + bkgd
        return dens, dens, dens, colors, zeros, zeros.clone() if t_starts is not None else None
    args = (_f32c(dens), _f32c(rgbs), pi, _f32c(t_starts), _f32c(t_ends), _f32c(prefix_trans), _f32c(bkgd),
            from_alpha, expected_depths, want_rays)
    if not torch.is_grad_enabled() or not (dens.requires_grad or (rgbs is not None and rgbs.requires_grad)):
        return _composite_forward(*args)[:6]  # inference: no graph node to build
    return _Composite.apply(*args)


# --------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------

def rendering(
    t_starts: Tensor,
    t_ends: Tensor,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    rgb_sigma_fn: Optional[Callable] = None,
    rgb_alpha_fn: Optional[Callable] = None,
    render_bkgd: Optional[Tensor] = None,
    expected_depths: bool = True,
) -> Tuple[Tensor, Tensor, Tensor, Dict]:
    """Render rays through the field given by `rgb_sigma_fn` / `rgb_alpha_fn`.

    Same contract as the reference (volrend.py:15-164): returns colors (n_rays, 3),
    opacities (n_rays, 1), depths (n_rays, 1) and an extras dict; differentiable
    w.r.t. what the closure returns, not w.r.t. t_starts / t_ends / ray_indices.
    """
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape, \
            "Since nerfacc 0.5.0, t_starts, t_ends and ray_indices must have the same shape (N,). "
    if rgb_sigma_fn is None and rgb_alpha_fn is None:
        raise ValueError("At least one of `rgb_sigma_fn` and `rgb_alpha_fn` should be specified.")

    use_sigma = rgb_sigma_fn is not None
    # the closure is called even for N == 0, as in the reference (volrend.py:91-96)
    rgbs, dens = (rgb_sigma_fn if use_sigma else rgb_alpha_fn)(t_starts, t_ends, ray_indices)
    assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
    name = "sigmas" if use_sigma else "alphas"
    assert dens.shape == t_starts.shape, "{} must have shape of (N,)! Got {}".format(name, dens.shape)

    t_grad = _grad_through(t_starts, t_ends)
    if ray_indices is not None and dens.is_cuda and not t_grad:
        assert n_rays is not None, "n_rays must be provided"
        bk = render_bkgd
        fuse_bkgd = bk is not None and not bk.requires_grad and bk.numel() == 3 and bk.is_cuda
        weights, trans, alphas, colors, opacities, depths = _composite_packed(
            dens, rgbs, t_starts, t_ends, None, ray_indices, n_rays, None,
            bk.reshape(3).to(torch.float32) if fuse_bkgd else None,
            from_alpha=not use_sigma, expected_depths=expected_depths, want_rays=True)
        if bk is not None and not fuse_bkgd:
            colors = colors + bk * (1.0 - opacities)
        extras = {"weights": weights, "alphas": dens if not use_sigma else alphas, "trans": trans, "rgbs": rgbs}
        if use_sigma:
            extras["sigmas"] = dens
        return colors, opacities, depths, extras

    seg = _batched_segments(dens, None) if ray_indices is None and rgbs.dtype == torch.float32 and not t_grad else None
    if seg is not None and t_starts.shape == dens.shape == t_ends.shape:
        # batched (n_rays, S) on the GPU: the same fused kernels, addressed through uniform segments
        bk = render_bkgd
        fuse_bkgd = bk is not None and not bk.requires_grad and bk.numel() == 3 and bk.is_cuda
        weights, trans, alphas, colors, opacities, depths = _composite_packed(
            dens.reshape(-1), rgbs.reshape(-1, 3), t_starts.reshape(-1), t_ends.reshape(-1), seg, None, None, None,
            bk.reshape(3).to(torch.float32) if fuse_bkgd else None,
            from_alpha=not use_sigma, expected_depths=expected_depths, want_rays=True)
        if bk is not None and not fuse_bkgd:
            colors = colors + bk * (1.0 - opacities)
        weights, trans = weights.view_as(dens), trans.view_as(dens)
        extras = {"weights": weights, "alphas": dens if not use_sigma else alphas.view_as(dens), "trans": trans,
                  "rgbs": rgbs}
        if use_sigma:
            extras["sigmas"] = dens
        return colors, opacities, depths, extras

    # batched CPU inputs, or gradients wanted w.r.t. t_starts / t_ends: the reference's own op sequence
    if use_sigma:
        weights, trans, alphas = render_weight_from_density(t_starts, t_ends, dens, ray_indices=ray_indices,
                                                            n_rays=n_rays)
        extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": dens, "rgbs": rgbs}
    else:
        weights, trans = render_weight_from_alpha(dens, ray_indices=ray_indices, n_rays=n_rays)
        extras = {"weights": weights, "trans": trans, "rgbs": rgbs, "alphas": dens}
    colors = accumulate_along_rays(weights, values=rgbs, ray_indices=ray_indices, n_rays=n_rays)
    opacities = accumulate_along_rays(weights, values=None, ray_indices=ray_indices, n_rays=n_rays)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0,
                                   ray_indices=ray_indices, n_rays=n_rays)
    if expected_depths:
        depths = depths / opacities.clamp_min(torch.finfo(rgbs.dtype).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    return colors, opacities, depths, extras


def render_transmittance_from_alpha(
    alphas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    prefix_trans: Optional[Tensor] = None,
) -> Tensor:
    """T_i = prod_{j<i} (1 - alpha_j) per ray (reference volrend.py:167-216)."""
    if _packed_mode(alphas, packed_info, ray_indices) and alphas.is_cuda and not _grad_through(prefix_trans):
        _, trans, _, _, _, _ = _composite_packed(alphas, None, None, None, packed_info, ray_indices, n_rays,
                                                 prefix_trans, None, True, False, False)
        return trans
    seg = None if _packed_mode(alphas, packed_info, ray_indices) else _batched_segments(alphas, prefix_trans)
    if seg is not None:
        return render_transmittance_from_alpha(alphas.reshape(-1), packed_info=seg).view_as(alphas)
    trans = exclusive_prod(1 - alphas, packed_info=packed_info, indices=ray_indices)
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans


def render_transmittance_from_density(
    t_starts: Tensor,
    t_ends: Tensor,
    sigmas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    prefix_trans: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """T_i = exp(-sum_{j<i} sigma_j dt_j) and alpha_i = 1 - exp(-sigma_i dt_i) (reference volrend.py:219-278)."""
    t_grad = _grad_through(t_starts, t_ends, prefix_trans)
    if _packed_mode(sigmas, packed_info, ray_indices) and sigmas.is_cuda and not t_grad:
        _, trans, alphas, _, _, _ = _composite_packed(sigmas, None, t_starts, t_ends, packed_info, ray_indices,
                                                      n_rays, prefix_trans, None, False, False, False)
        return trans, alphas
    seg = None if (t_grad or _packed_mode(sigmas, packed_info, ray_indices)) else _batched_segments(sigmas, prefix_trans)
    if seg is not None and t_starts.shape == sigmas.shape == t_ends.shape:
        trans, alphas = render_transmittance_from_density(t_starts.reshape(-1), t_ends.reshape(-1),
                                                          sigmas.reshape(-1), packed_info=seg)
        return trans.view_as(sigmas), alphas.view_as(sigmas)
    sigmas_dt = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sigmas_dt)
    trans = torch.exp(-exclusive_sum(sigmas_dt, packed_info=packed_info, indices=ray_indices))
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans, alphas


def render_weight_from_alpha(
    alphas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    prefix_trans: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """w_i = T_i alpha_i (reference volrend.py:281-323).  Returns (weights, trans)."""
    if _packed_mode(alphas, packed_info, ray_indices) and alphas.is_cuda and not _grad_through(prefix_trans):
        weights, trans, _, _, _, _ = _composite_packed(alphas, None, None, None, packed_info, ray_indices, n_rays,
                                                       prefix_trans, None, True, False, False)
        return weights, trans
    seg = None if _packed_mode(alphas, packed_info, ray_indices) else _batched_segments(alphas, prefix_trans)
    if seg is not None:
        weights, trans = render_weight_from_alpha(alphas.reshape(-1), packed_info=seg)
        return weights.view_as(alphas), trans.view_as(alphas)
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    return trans * alphas, trans


def render_weight_from_density(
    t_starts: Tensor,
    t_ends: Tensor,
    sigmas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    prefix_trans: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """w_i = T_i (1 - exp(-sigma_i dt_i)) (reference volrend.py:326-376).  Returns (weights, trans, alphas)."""
    t_grad = _grad_through(t_starts, t_ends, prefix_trans)
    if _packed_mode(sigmas, packed_info, ray_indices) and sigmas.is_cuda and not t_grad:
        weights, trans, alphas, _, _, _ = _composite_packed(sigmas, None, t_starts, t_ends, packed_info,
                                                            ray_indices, n_rays, prefix_trans, None, False, False,
                                                            False)
        return weights, trans, alphas
    seg = None if (t_grad or _packed_mode(sigmas, packed_info, ray_indices)) else _batched_segments(sigmas, prefix_trans)
    if seg is not None and t_starts.shape == sigmas.shape == t_ends.shape:
        weights, trans, alphas = render_weight_from_density(t_starts.reshape(-1), t_ends.reshape(-1),
                                                            sigmas.reshape(-1), packed_info=seg)
        return weights.view_as(sigmas), trans.view_as(sigmas), alphas.view_as(sigmas)
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                      prefix_trans)
    return trans * alphas, trans, alphas


@torch.no_grad()
def render_visibility_from_alpha(
    alphas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    early_stop_eps: float = 1e-4,
    alpha_thre: float = 0.0,
    prefix_trans: Optional[Tensor] = None,
) -> Tensor:
    """Boolean mask: transmittance >= early_stop_eps and (alpha >= alpha_thre) (reference volrend.py:379-432)."""
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


@torch.no_grad()
def render_visibility_from_density(
    t_starts: Tensor,
    t_ends: Tensor,
    sigmas: Tensor,
    packed_info: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
    early_stop_eps: float = 1e-4,
    alpha_thre: float = 0.0,
    prefix_trans: Optional[Tensor] = None,
) -> Tensor:
    """Same as :func:`render_visibility_from_alpha` from densities (reference volrend.py:435-494)."""
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                      prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def accumulate_along_rays(
    weights: Tensor,
    values: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    n_rays: Optional[int] = None,
) -> Tensor:
    """out[r] = sum over the samples of ray r of w_i * v_i (reference volrend.py:497-561).

    `values=None` accumulates the weights (D == 1).  Differentiable w.r.t. weights and values.
    """
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    if ray_indices is None:
        src = weights[..., None] if values is None else weights[..., None] * values
        return torch.sum(src, dim=-2)
    assert n_rays is not None, "n_rays must be provided"
    assert weights.dim() == 1, "weights must be flattened"
    if not weights.is_cuda:
        src = weights[..., None] if values is None else weights[..., None] * values
        out = torch.zeros((n_rays, src.shape[-1]), device=src.device, dtype=src.dtype)
        out.index_add_(0, ray_indices, src)
        return out
    idx = ray_indices.contiguous()
    if idx.dtype != torch.int64:
        idx = idx.to(torch.int64)
    # segments are only trusted when they came out of the traversal together with these indices
    pi = _stashed_packed_info(ray_indices, n_rays)
    return _Accumulate.apply(_f32c(weights), _f32c(values), idx, pi, int(n_rays))


def accumulate_along_rays_(
    weights: Tensor,
    values: Optional[Tensor] = None,
    ray_indices: Optional[Tensor] = None,
    outputs: Optional[Tensor] = None,
) -> None:
    """In-place variant: adds into `outputs` (reference volrend.py:564-587)."""
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    dim = 1 if values is None else values.shape[-1]
    if ray_indices is None:
        src = weights[..., None] if values is None else weights[..., None] * values
        outputs.add_(src.sum(dim=-2))
        return
    assert weights.dim() == 1, "weights must be flattened"
    assert outputs.dim() == 2 and outputs.shape[-1] == dim, "outputs must be of shape (n_rays, D)"
    if (not weights.is_cuda) or outputs.dtype != torch.float32 or not outputs.is_contiguous():
        src = weights[..., None] if values is None else weights[..., None] * values
        outputs.index_add_(0, ray_indices, src)
        return
    if weights.shape[0] == 0:
        return
    idx = ray_indices.contiguous()
    if idx.dtype != torch.int64:
        idx = idx.to(torch.int64)
    w, v = _f32c(weights.detach()), _f32c(None if values is None else values.detach())
    _lib.call("nfa_accumulate_atomic", weights.device, w.shape[0], _lib.ptr(idx), _lib.ptr(w), _lib.ptr(v), dim,
              _lib.ptr(outputs))
