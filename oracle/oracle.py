"""numpy front-end of the CPU oracle (oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Imported by tests/,
``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs; never by the
``nerfacc_b200`` package.

The function names and argument meaning follow the reference's Python layer
(/root/reference/nerfacc/{grid,pack,scan,volrend}.py and
estimators/occ_grid.py) so that tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.orc_traverse_ctx_size() == C.sizeof(_TraverseCtx)
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """Override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_num_threads(int(n))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_U8 = C.POINTER(C.c_uint8)


class _TraverseCtx(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int32),
        ("rays_o", _F), ("rays_d", _F), ("rays_mask", _U8),
        ("n_grids", C.c_int32), ("res", C.c_int32 * 3),
        ("binaries", _U8), ("aabbs", _F),
        ("hits", _U8), ("t_sorted", _F), ("t_indices", _I64),
        ("near_planes", _F), ("far_planes", _F),
        ("step_size", C.c_float), ("cone_angle", C.c_float), ("steps_limit", C.c_int32),
        ("iv_starts", _I64), ("iv_cnts", _I64), ("iv_vals", _F), ("iv_ray", _I64),
        ("iv_left", _U8), ("iv_right", _U8),
        ("sm_starts", _I64), ("sm_cnts", _I64), ("sm_vals", _F), ("sm_ray", _I64), ("sm_valid", _U8),
        ("terminate_planes", _F),
    ]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _u8(a):
    return np.ascontiguousarray(np.asarray(a).astype(np.uint8))


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


# --------------------------------------------------------------------------
# grid.py
# --------------------------------------------------------------------------

def ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane=-np.inf, far_plane=np.inf, miss_value=np.inf):
    """reference: nerfacc/grid.py:13-51 -> csrc/grid.cu:284-313."""
    rays_o, rays_d, aabbs = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    R, G = rays_o.shape[0], aabbs.shape[0]
    t_mins = np.empty((R, G), np.float32)
    t_maxs = np.empty((R, G), np.float32)
    hits = np.empty((R, G), np.uint8)
    lib().orc_ray_aabb_intersect(
        C.c_int32(R), _p(rays_o, _F), _p(rays_d, _F), C.c_int32(G), _p(aabbs, _F),
        C.c_float(near_plane), C.c_float(far_plane), C.c_float(miss_value),
        _p(t_mins, _F), _p(t_maxs, _F), _p(hits, _U8))
    return t_mins, t_maxs, hits.astype(bool)


def sort_intersections(t_mins, t_maxs):
    """reference: nerfacc/grid.py:160-162 (torch.sort of cat([t_mins, t_maxs]))."""
    t_mins, t_maxs = _f32(t_mins), _f32(t_maxs)
    R, G = t_mins.shape
    t_sorted = np.empty((R, 2 * G), np.float32)
    t_indices = np.empty((R, 2 * G), np.int64)
    lib().orc_sort_intersections(C.c_int32(R), C.c_int32(G), _p(t_mins, _F), _p(t_maxs, _F),
                                 _p(t_sorted, _F), _p(t_indices, _I64))
    return t_sorted, t_indices


def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=None, far_planes=None,
                   step_size=1e-3, cone_angle=0.0, traverse_steps_limit=None, over_allocate=False,
                   rays_mask=None, t_sorted=None, t_indices=None, hits=None):
    """reference: nerfacc/grid.py:93-192 -> csrc/grid.cu:320-474.

    Returns (intervals, samples, terminate_planes) with the reference's field
    names: intervals = dict(vals, ray_indices, is_left, is_right, packed_info),
    samples = dict(vals, ray_indices, is_valid, packed_info).
    """
    rays_o, rays_d, aabbs = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    binaries = _u8(binaries)
    R, G = rays_o.shape[0], binaries.shape[0]
    near_planes = np.zeros(R, np.float32) if near_planes is None else _f32(near_planes)
    far_planes = np.full(R, np.inf, np.float32) if far_planes is None else _f32(far_planes)
    mask = np.ones(R, np.uint8) if rays_mask is None else _u8(rays_mask)
    limit = -1 if traverse_steps_limit is None else int(traverse_steps_limit)
    if over_allocate:
        assert limit > 0
    if t_sorted is None or t_indices is None or hits is None:
        t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
        t_sorted, t_indices = sort_intersections(t_mins, t_maxs)
    t_sorted, t_indices, hits = _f32(t_sorted), _i64(t_indices), _u8(hits)

    ctx = _TraverseCtx()
    ctx.n_rays, ctx.n_grids = R, G
    ctx.res = (C.c_int32 * 3)(*binaries.shape[1:])
    ctx.rays_o, ctx.rays_d = _p(rays_o, _F), _p(rays_d, _F)
    ctx.binaries, ctx.aabbs = _p(binaries, _U8), _p(aabbs, _F)
    ctx.hits, ctx.t_sorted, ctx.t_indices = _p(hits, _U8), _p(t_sorted, _F), _p(t_indices, _I64)
    ctx.near_planes, ctx.far_planes = _p(near_planes, _F), _p(far_planes, _F)
    ctx.step_size, ctx.cone_angle, ctx.steps_limit = step_size, cone_angle, limit
    # two-pass mode never writes the plane of a ray without samples (grid.cu:103-106
    # skips it in the fill pass and the buffer comes from torch::empty): NaN marks "undefined".
    term = np.full(R, np.nan, np.float32)
    iv_cnts = np.zeros(R, np.int64)
    sm_cnts = np.zeros(R, np.int64)
    keep = [iv_cnts, sm_cnts, term, mask]

    def alloc(cnts, edges):
        starts = np.cumsum(cnts) - cnts
        n = int(cnts.sum())
        return starts.astype(np.int64), n

    if over_allocate:
        # csrc/grid.cu:364-404: fixed stride per (masked-in) ray, single pass.
        iv_cnts[:] = 2 * limit * mask.astype(np.int64)
        sm_cnts[:] = limit * mask.astype(np.int64)
        ctx.rays_mask = _p(mask, _U8)
    else:
        # csrc/grid.cu:405-470: the mask is NOT passed in two-pass mode.
        ctx.rays_mask = None
        ctx.iv_cnts, ctx.sm_cnts = _p(iv_cnts, _I64), _p(sm_cnts, _I64)
        lib().orc_traverse_pass(C.byref(ctx), C.c_int32(0))

    iv_starts, n_iv = alloc(iv_cnts, True)
    sm_starts, n_sm = alloc(sm_cnts, False)
    iv_vals = np.zeros(n_iv, np.float32); iv_ray = np.zeros(n_iv, np.int64)
    iv_left = np.zeros(n_iv, np.uint8); iv_right = np.zeros(n_iv, np.uint8)
    sm_vals = np.zeros(n_sm, np.float32); sm_ray = np.zeros(n_sm, np.int64)
    sm_valid = np.zeros(n_sm, np.uint8)
    keep += [iv_starts, sm_starts]
    ctx.iv_starts, ctx.iv_cnts = _p(iv_starts, _I64), _p(iv_cnts, _I64)
    ctx.sm_starts, ctx.sm_cnts = _p(sm_starts, _I64), _p(sm_cnts, _I64)
    ctx.iv_vals, ctx.iv_ray = _p(iv_vals, _F), _p(iv_ray, _I64)
    ctx.iv_left, ctx.iv_right = _p(iv_left, _U8), _p(iv_right, _U8)
    ctx.sm_vals, ctx.sm_ray, ctx.sm_valid = _p(sm_vals, _F), _p(sm_ray, _I64), _p(sm_valid, _U8)
    ctx.terminate_planes = _p(term, _F)
    lib().orc_traverse_pass(C.byref(ctx), C.c_int32(1))
    if over_allocate:
        # grid.cu:402-404: chunk_starts recomputed from the actual counts
        iv_starts = (np.cumsum(iv_cnts) - iv_cnts).astype(np.int64)
        sm_starts = (np.cumsum(sm_cnts) - sm_cnts).astype(np.int64)
    intervals = dict(vals=iv_vals, ray_indices=iv_ray, is_left=iv_left.astype(bool),
                     is_right=iv_right.astype(bool), packed_info=np.stack([iv_starts, iv_cnts], -1))
    samples = dict(vals=sm_vals, ray_indices=sm_ray, is_valid=sm_valid.astype(bool),
                   packed_info=np.stack([sm_starts, sm_cnts], -1))
    return intervals, samples, term


def occgrid_sampling(rays_o, rays_d, binaries, aabbs, near_plane=0.0, far_plane=1e10,
                     t_min=None, t_max=None, render_step_size=1e-3, cone_angle=0.0,
                     near_jitter=None):
    """reference: nerfacc/estimators/occ_grid.py:154-177 (no sigma_fn / alpha_fn).

    ``near_jitter`` stands in for ``torch.rand_like(near_planes)`` of the
    stratified branch (occ_grid.py:162-163): pass the [n_rays] uniform draws.
    Returns (ray_indices, t_starts, t_ends, packed_info).
    """
    rays_o = _f32(rays_o)
    R = rays_o.shape[0]
    near = np.full(R, near_plane, np.float32)
    far = np.full(R, far_plane, np.float32)
    if t_min is not None:
        near = np.maximum(near, _f32(t_min))
    if t_max is not None:
        far = np.minimum(far, _f32(t_max))
    if near_jitter is not None:
        near = near + _f32(near_jitter) * np.float32(render_step_size)
    iv, sm, _ = traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=near, far_planes=far,
                               step_size=render_step_size, cone_angle=cone_angle)
    t_starts = iv["vals"][iv["is_left"]]
    t_ends = iv["vals"][iv["is_right"]]
    return sm["ray_indices"], t_starts, t_ends, sm["packed_info"]


# --------------------------------------------------------------------------
# pack.py / scan.py
# --------------------------------------------------------------------------

def pack_info(ray_indices, n_rays: Optional[int] = None):
    """reference: nerfacc/pack.py:10-49."""
    ray_indices = _i64(ray_indices)
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.size else 0
    out = np.zeros((n_rays, 2), np.int64)
    lib().orc_pack_info(C.c_int64(ray_indices.size), _p(ray_indices, _I64), C.c_int32(n_rays), _p(out, _I64))
    return out


def _scan(inputs, packed_info, indices, op, inclusive, reverse=False, normalize=False):
    inputs = _f32(inputs)
    out = np.empty_like(inputs)
    if packed_info is not None and indices is not None:
        raise ValueError("Only one of `indices` and `packed_info` can be specified.")
    if packed_info is not None:
        pi = _i64(packed_info)
        starts, cnts = np.ascontiguousarray(pi[:, 0]), np.ascontiguousarray(pi[:, 1])
        lib().orc_scan_packed(C.c_int32(pi.shape[0]), _p(starts, _I64), _p(cnts, _I64), _p(inputs, _F),
                              _p(out, _F), C.c_int32(op), C.c_int32(inclusive), C.c_int32(reverse),
                              C.c_int32(normalize))
    elif indices is not None:
        keys = _i64(indices)
        lib().orc_scan_by_key(C.c_int64(inputs.size), _p(keys, _I64), _p(inputs, _F), _p(out, _F),
                              C.c_int32(op), C.c_int32(inclusive), C.c_int32(reverse))
    else:
        flat = inputs.reshape(-1, inputs.shape[-1])
        R, S = flat.shape
        starts = (np.arange(R, dtype=np.int64) * S)
        cnts = np.full(R, S, np.int64)
        o = np.empty_like(flat)
        lib().orc_scan_packed(C.c_int32(R), _p(starts, _I64), _p(cnts, _I64), _p(flat, _F), _p(o, _F),
                              C.c_int32(op), C.c_int32(inclusive), C.c_int32(reverse), C.c_int32(normalize))
        out = o.reshape(inputs.shape)
    return out


def inclusive_sum(inputs, packed_info=None, indices=None, reverse=False):
    """reference: nerfacc/scan.py:14-77."""
    return _scan(inputs, packed_info, indices, 0, 1, reverse)


def exclusive_sum(inputs, packed_info=None, indices=None, reverse=False):
    """reference: nerfacc/scan.py:80-145."""
    return _scan(inputs, packed_info, indices, 0, 0, reverse)


def inclusive_prod(inputs, packed_info=None, indices=None):
    """reference: nerfacc/scan.py:148-211."""
    return _scan(inputs, packed_info, indices, 1, 1)


def exclusive_prod(inputs, packed_info=None, indices=None):
    """reference: nerfacc/scan.py:214-282."""
    return _scan(inputs, packed_info, indices, 1, 0)


def prod_backward(inputs, outputs, grad_outputs, packed_info=None, indices=None, inclusive=True):
    """reference: csrc/scan.cu:199-210,289-300 / scan_cub.cu:205-211,274-280:
    reverse scan of grad*out (inclusive for inclusive_prod, exclusive for
    exclusive_prod), then divided by inputs.clamp_min(1e-10)."""
    g = _f32(grad_outputs) * _f32(outputs)
    s = _scan(g, packed_info, indices, 0, 1 if inclusive else 0, reverse=True)
    return s / np.maximum(_f32(inputs), np.float32(1e-10))


# --------------------------------------------------------------------------
# volrend.py
# --------------------------------------------------------------------------

def _starts_cnts(packed_info, ray_indices, n_rays):
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    pi = _i64(packed_info)
    return np.ascontiguousarray(pi[:, 0]), np.ascontiguousarray(pi[:, 1])


def composite(t_starts, t_ends, sigmas, rgbs=None, packed_info=None, ray_indices=None, n_rays=None,
              prefix_trans=None, render_bkgd=None, expected_depths=True):
    """Packed ``rendering`` forward, density route.
    reference: nerfacc/volrend.py:79-164,219-278,326-376,497-561.
    Returns dict(weights, trans, alphas, colors, opacities, depths)."""
    t_starts, t_ends, sigmas = _f32(t_starts), _f32(t_ends), _f32(sigmas)
    starts, cnts = _starts_cnts(packed_info, ray_indices, n_rays)
    R, N = starts.shape[0], t_starts.shape[0]
    rgbs = None if rgbs is None else _f32(rgbs)
    pt = None if prefix_trans is None else _f32(prefix_trans)
    bg = None if render_bkgd is None else _f32(render_bkgd)
    w = np.empty(N, np.float32); T = np.empty(N, np.float32); a = np.empty(N, np.float32)
    col = np.zeros((R, 3), np.float32) if rgbs is not None else None
    op = np.zeros((R, 1), np.float32); dep = np.zeros((R, 1), np.float32)
    lib().orc_composite_fwd(C.c_int32(R), _p(starts, _I64), _p(cnts, _I64), _p(t_starts, _F), _p(t_ends, _F),
                            _p(sigmas, _F), _p(rgbs, _F), _p(pt, _F), _p(bg, _F), C.c_int32(bool(expected_depths)),
                            _p(w, _F), _p(T, _F), _p(a, _F), _p(col, _F), _p(op, _F), _p(dep, _F))
    return dict(weights=w, trans=T, alphas=a, colors=col, opacities=op, depths=dep)


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans=None):
    """reference: nerfacc/volrend.py:326-376."""
    o = composite(t_starts, t_ends, sigmas, None, packed_info, ray_indices, n_rays, prefix_trans)
    return o["weights"], o["trans"], o["alphas"]


def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """reference: nerfacc/volrend.py:281-323."""
    alphas = _f32(alphas)
    starts, cnts = _starts_cnts(packed_info, ray_indices, n_rays)
    pt = None if prefix_trans is None else _f32(prefix_trans)
    w = np.empty_like(alphas); T = np.empty_like(alphas)
    lib().orc_composite_alpha_fwd(C.c_int32(starts.shape[0]), _p(starts, _I64), _p(cnts, _I64), _p(alphas, _F),
                                  _p(pt, _F), _p(w, _F), _p(T, _F))
    return w, T


def composite_backward(t_starts, t_ends, sigmas, rgbs, packed_info, gC=None, gO=None, gD=None,
                       gW=None, gT=None, gA=None, prefix_trans=None, render_bkgd=None, expected_depths=True):
    """float64 gradient of a scalar loss through :func:`composite`.
    Returns (g_sigmas, g_rgbs) as float64."""
    t_starts, t_ends, sigmas = _f32(t_starts), _f32(t_ends), _f32(sigmas)
    starts, cnts = _starts_cnts(packed_info, None, None)
    R, N = starts.shape[0], t_starts.shape[0]
    rgbs = None if rgbs is None else _f32(rgbs)
    conv = lambda x: None if x is None else _f32(x)
    gC, gO, gD, gW, gT, gA = map(conv, (gC, gO, gD, gW, gT, gA))
    pt, bg = conv(prefix_trans), conv(render_bkgd)
    gs = np.zeros(N, np.float64)
    gr = np.zeros((N, 3), np.float64) if rgbs is not None else None
    lib().orc_composite_bwd(C.c_int32(R), _p(starts, _I64), _p(cnts, _I64), _p(t_starts, _F), _p(t_ends, _F),
                            _p(sigmas, _F), _p(rgbs, _F), _p(pt, _F), _p(bg, _F), C.c_int32(bool(expected_depths)),
                            _p(gC, _F), _p(gO, _F), _p(gD, _F), _p(gW, _F), _p(gT, _F), _p(gA, _F),
                            _p(gs, _D), _p(gr, _D))
    return gs, gr


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """reference: nerfacc/volrend.py:497-561 (packed branch)."""
    weights = _f32(weights)
    ray_indices = _i64(ray_indices)
    values = None if values is None else _f32(values)
    dim = 1 if values is None else values.shape[-1]
    out = np.zeros((n_rays, dim), np.float32)
    lib().orc_accumulate(C.c_int32(n_rays), C.c_int64(weights.size), _p(ray_indices, _I64), _p(weights, _F),
                         _p(values, _F), C.c_int32(dim), _p(out, _F))
    return out


# --------------------------------------------------------------------------
# pdf.py
# --------------------------------------------------------------------------

_U8P = C.POINTER(C.c_uint8)


def philox_uniform(seed: int, subsequence: int, offset: int) -> float:
    """First curand_uniform() after curand_init(seed, subsequence, offset) -- pdf.cu:139-145."""
    fn = lib().orc_philox_uniform
    fn.restype = C.c_float
    return float(fn(C.c_uint64(seed), C.c_uint64(subsequence), C.c_uint64(offset)))


def importance_sampling(vals, cdfs, n_intervals_per_ray, packed_info=None, stratified=False, seed=0, offset=0):
    """reference: nerfacc/pdf.py:64-131 -> pdf.cu:293-426.

    ``vals`` / ``cdfs``: batched [n_rays, E] (packed_info None) or flattened [all_edges] with ``packed_info``.
    ``n_intervals_per_ray``: int -> batched outputs ([n_rays, n+1] edges, [n_rays, n] centres);
    array [n_rays] -> flattened outputs, returned as two dicts (vals, packed_info, ray_indices, is_left, is_right).
    """
    vals, cdfs = _f32(vals), _f32(cdfs)
    if packed_info is None:
        assert vals.ndim >= 2
        n_rays, in_edges = int(np.prod(vals.shape[:-1])), vals.shape[-1]
        lead = vals.shape[:-1]
        pin = None
    else:
        pin = _i64(packed_info)
        n_rays, in_edges, lead = pin.shape[0], 0, (pin.shape[0],)
    fn = lib().orc_importance_sampling
    if np.ndim(n_intervals_per_ray) == 0:
        n = int(n_intervals_per_ray)
        s_vals = np.zeros(lead + (n,), np.float32)
        e_vals = np.zeros(lead + (n + 1,), np.float32)
        fn(C.c_int32(n_rays), _p(vals, _F), _p(cdfs, _F), _p(pin, _I64), C.c_int64(in_edges), None, None,
           C.c_int64(n), C.c_int32(bool(stratified)), C.c_uint64(seed), C.c_uint64(offset), _p(s_vals, _F), None,
           _p(e_vals, _F), None, None, None)
        return e_vals, s_vals
    cnts = _i64(n_intervals_per_ray).reshape(-1)
    assert cnts.size == n_rays
    s_pack = np.stack([np.cumsum(cnts) - cnts, cnts], -1).astype(np.int64)
    e_cnt = (cnts + 1) * (cnts > 0)
    e_pack = np.stack([np.cumsum(e_cnt) - e_cnt, e_cnt], -1).astype(np.int64)
    ns, ne = int(cnts.sum()), int(e_cnt.sum())
    s_vals, s_ray = np.zeros(ns, np.float32), np.zeros(ns, np.int64)
    e_vals, e_ray = np.zeros(ne, np.float32), np.zeros(ne, np.int64)
    e_left, e_right = np.zeros(ne, np.uint8), np.zeros(ne, np.uint8)
    fn(C.c_int32(n_rays), _p(vals, _F), _p(cdfs, _F), _p(pin, _I64), C.c_int64(in_edges), _p(s_pack, _I64),
       _p(e_pack, _I64), C.c_int64(0), C.c_int32(bool(stratified)), C.c_uint64(seed), C.c_uint64(offset),
       _p(s_vals, _F), _p(s_ray, _I64), _p(e_vals, _F), _p(e_ray, _I64), _p(e_left, _U8P), _p(e_right, _U8P))
    return (dict(vals=e_vals, packed_info=e_pack, ray_indices=e_ray, is_left=e_left.astype(bool),
                 is_right=e_right.astype(bool)),
            dict(vals=s_vals, packed_info=s_pack, ray_indices=s_ray))


def searchsorted(key_vals, query_vals, key_packed_info=None, query_packed_info=None, query_ray_indices=None):
    """reference: nerfacc/pdf.py:12-61 -> pdf.cu:429-456.  Returns (ids_left, ids_right)."""
    kv, qv = _f32(key_vals), _f32(query_vals)
    kp = None if key_packed_info is None else _i64(key_packed_info)
    qp = None if query_packed_info is None else _i64(query_packed_info)
    qr = None if query_ray_indices is None else _i64(query_ray_indices)
    n_rays = qp.shape[0] if qp is not None else int(np.prod(qv.shape[:-1]))
    left, right = np.zeros(qv.shape, np.int64), np.zeros(qv.shape, np.int64)
    lib().orc_searchsorted(C.c_int64(qv.size), _p(qv, _F), _p(qp, _I64), _p(qr, _I64), C.c_int32(n_rays),
                           C.c_int64(qv.shape[-1] if qp is None else 0), _p(kv, _F), _p(kp, _I64),
                           C.c_int64(kv.shape[-1] if kp is None else 0), _p(left, _I64), _p(right, _I64))
    return left, right


# --------------------------------------------------------------------------
# estimators/occ_grid.py: grid maintenance
# --------------------------------------------------------------------------

def occ_ema_update(occs, cell_ids, occ, ema_decay=0.95):
    """occs[cell_ids] = maximum(occs[cell_ids] * ema_decay, occ) (occ_grid.py:395-398); returns the new array."""
    out = _f32(occs).copy()
    ids, val = _i64(cell_ids), _f32(occ)
    fresh = np.empty(len(ids), np.float32)
    lib().orc_occ_ema_update(C.c_int64(len(ids)), _p(ids, _I64), _p(val, _F), C.c_float(ema_decay), _p(out, _F), _p(fresh, _F))
    return out


def occ_threshold(occs, occ_thre=0.01):
    """(binaries, thre) with thre = min(mean(occs[occs >= 0]), occ_thre), binaries = occs > thre (occ_grid.py:400-404)."""
    o = _f32(occs)
    b = np.empty(o.size, np.uint8)
    fn = lib().orc_occ_threshold
    fn.restype = C.c_float
    thre = fn(C.c_int64(o.size), _p(o, _F), C.c_float(occ_thre), _p(b, _U8))
    return b.astype(bool).reshape(o.shape), float(thre)
