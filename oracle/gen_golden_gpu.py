"""Generate golden vectors from the UNMODIFIED reference CUDA build.

Run on a GPU box from the repo root:

    python oracle/gen_golden_gpu.py            # writes gpurun_out/golden/*.npz

The reference is imported from ``baseline/_ref`` (``pip install --target
baseline/_ref /root/reference``); nothing from nerfacc_b200 is
imported here.  The resulting files are committed under ``tests/golden/`` and pin
both the oracle (tests/test_oracle_golden.py, CPU) and the native path
(tests/test_gpu_golden.py).  Inputs are stored next to the outputs so the tests
do not depend on RNG stream stability.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import nerfacc as ref  # noqa: E402  (the reference)

assert "baseline/_ref" in ref.__file__.replace("\\", "/"), ref.__file__

# scene helpers are plain numpy; import the module file directly to avoid importing the package
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "nerfacc_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(scenes)

dev = torch.device("cuda:0")
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in arrays.items() if hasattr(v, "shape")}, os.path.getsize(path), "bytes")


def estimator(bins, aabbs):
    est = ref.OccGridEstimator(torch.from_numpy(aabbs[0]), resolution=list(bins.shape[1:]), levels=bins.shape[0]).to(dev)
    est.binaries = T(bins)
    return est


def sampling_case(name, ro, rd, bins, aabbs, **kw):
    est = estimator(bins, aabbs)
    tk = {k: (T(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    ri, ts, te = est.sampling(T(ro), T(rd), **tk)
    n_rays = ro.shape[0]
    pi = N(ref.pack_info(ri, n_rays))
    extra = {("in_" + k): v for k, v in kw.items() if isinstance(v, np.ndarray)}
    scal = {k: v for k, v in kw.items() if not isinstance(v, np.ndarray)}
    save(name, rays_o=ro, rays_d=rd, binaries_bits=np.packbits(bins.astype(np.uint8).ravel()),
         binaries_shape=np.array(bins.shape), aabbs=aabbs, kw_names=np.array(sorted(scal)),
         kw_vals=np.array([float(scal[k]) for k in sorted(scal)], dtype=np.float64),
         packed_info=pi, t_starts=N(ts), t_ends=N(te), ray_indices_sha=np.array(sha(N(ri))),
         n_samples=np.array(ri.numel()), **extra)


rng = np.random.default_rng(2024)

# (a) ball scene, the bench geometry at small R
ro, rd = scenes.ball_rays(1024, seed=42)
ball = scenes.ball_grid(128)
sampling_case("ref_sampling_ball", ro, rd, ball, scenes.nested_aabbs(1), render_step_size=scenes.BALL_STEP)

# (b) fragmented ball: many runs per ray
frag = ball & (rng.random(ball.shape) > 0.5)
sampling_case("ref_sampling_frag", ro[:512], rd[:512], frag, scenes.nested_aabbs(1), render_step_size=scenes.BALL_STEP)

# (c) stratified-style per-ray near planes + per-ray t_min / t_max on 4 nested random levels
R2 = 96
ro2 = rng.standard_normal((R2, 3)).astype(np.float32)
rd2 = rng.standard_normal((R2, 3)).astype(np.float32)
rd2 /= np.linalg.norm(rd2, axis=1, keepdims=True)
bins4 = rng.random((4, 32, 32, 32)) > 0.5
tmin = rng.random(R2).astype(np.float32)
tmax = (tmin + 3 * rng.random(R2)).astype(np.float32)
sampling_case("ref_sampling_lvl4_tminmax", ro2, rd2, bins4, scenes.nested_aabbs(4), render_step_size=1e-2,
              near_plane=0.15, far_plane=3.4, t_min=tmin, t_max=tmax)
sampling_case("ref_sampling_lvl4", ro2, rd2, bins4, scenes.nested_aabbs(4), render_step_size=7e-3)

# (d) traverse_grids full outputs (edges, flags, midpoints, terminate planes)
iv, sm, term = ref.traverse_grids(T(ro2), T(rd2), T(bins4), T(scenes.nested_aabbs(4)), step_size=1e-2)
save("ref_traverse_lvl4", rays_o=ro2, rays_d=rd2, binaries_bits=np.packbits(bins4.astype(np.uint8).ravel()),
     binaries_shape=np.array(bins4.shape), aabbs=scenes.nested_aabbs(4), step_size=np.array(1e-2),
     iv_vals=N(iv.vals), iv_left=np.packbits(N(iv.is_left)), iv_right=np.packbits(N(iv.is_right)),
     iv_packed_info=N(iv.packed_info), iv_ray_sha=np.array(sha(N(iv.ray_indices))),
     sm_vals=N(sm.vals), sm_packed_info=N(sm.packed_info), sm_ray_sha=np.array(sha(N(sm.ray_indices))),
     terminate=N(term))

# (e) ray_aabb_intersect on the reference test's shapes
ro3 = rng.random((200, 3)).astype(np.float32)
rd3 = rng.standard_normal((200, 3)).astype(np.float32)
rd3 /= np.linalg.norm(rd3, axis=1, keepdims=True)
bmin = rng.random((20, 3)).astype(np.float32)
boxes = np.concatenate([bmin, bmin + rng.random((20, 3)).astype(np.float32)], -1)
tm, tM, h = ref.ray_aabb_intersect(T(ro3), T(rd3), T(boxes))
save("ref_ray_aabb", rays_o=ro3, rays_d=rd3, aabbs=boxes, t_mins=N(tm), t_maxs=N(tM), hits=N(h))

# (f) rendering forward + backward on the ball samples
R = 96
est = estimator(ball, scenes.nested_aabbs(1))
ri, ts, te = est.sampling(T(ro[:R]), T(rd[:R]), render_step_size=scenes.BALL_STEP)
n = ri.numel()
sig_np = (5 * rng.random(n)).astype(np.float32)
rgb_np = rng.random((n, 3)).astype(np.float32)
sig = T(sig_np).requires_grad_(True)
rgb = T(rgb_np).requires_grad_(True)
bk = np.array([0.2, 0.5, 0.9], np.float32)
col, op, dep, ex = ref.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig), render_bkgd=T(bk))
gC = rng.random((R, 3)).astype(np.float32)
gO = rng.random((R, 1)).astype(np.float32)
gD = rng.random((R, 1)).astype(np.float32)
((col * T(gC)).sum() + (op * T(gO)).sum() + (dep * T(gD)).sum()).backward()
save("ref_render_ball", n_rays=np.array(R), packed_info=N(ref.pack_info(ri, R)), t_starts=N(ts), t_ends=N(te),
     sigmas=sig_np, rgbs=rgb_np, bkgd=bk, gC=gC, gO=gO, gD=gD,
     colors=N(col), opacities=N(op), depths=N(dep), weights=N(ex["weights"]), trans=N(ex["trans"]),
     alphas=N(ex["alphas"]), g_sigmas=N(sig.grad), g_rgbs=N(rgb.grad))

# alpha route + extras gradients
sig.grad = None
al = T((rng.random(n) * 0.3).astype(np.float32)).requires_grad_(True)
w, tr = ref.render_weight_from_alpha(al, ray_indices=ri, n_rays=R)
gW = rng.random(n).astype(np.float32)
gT = rng.random(n).astype(np.float32)
((w * T(gW)).sum() + (tr * T(gT)).sum()).backward()
w2, tr2, a2 = ref.render_weight_from_density(ts, te, sig, ray_indices=ri, n_rays=R)
gA = rng.random(n).astype(np.float32)
((w2 * T(gW)).sum() + (tr2 * T(gT)).sum() + (a2 * T(gA)).sum()).backward()
save("ref_render_extras", alphas_in=N(al), gW=gW, gT=gT, gA=gA, weights_a=N(w), trans_a=N(tr), g_alphas=N(al.grad),
     g_sigmas=N(sig.grad))

# (g) scans
cnts = rng.integers(0, 100, 120)
starts = np.cumsum(cnts) - cnts
pinfo = np.stack([starts, cnts], -1).astype(np.int64)
m = int(cnts.sum())
idx = np.repeat(np.arange(120), cnts).astype(np.int64)
x_np = (rng.random(m) * 0.2 + 0.9).astype(np.float32)
out = {}
for nm in ["inclusive_sum", "exclusive_sum", "inclusive_prod", "exclusive_prod"]:
    for mode in ["packed", "key"]:
        x = T(x_np).requires_grad_(True)
        y = getattr(ref, nm)(x, packed_info=T(pinfo)) if mode == "packed" else getattr(ref, nm)(x, indices=T(idx))
        gy = T((np.arange(m) % 7 + 1).astype(np.float32))
        (y * gy).sum().backward()
        out[f"{nm}_{mode}"] = N(y)
        out[f"{nm}_{mode}_grad"] = N(x.grad)
save("ref_scans", packed_info=pinfo, x=x_np, **out)
print("done ->", OUT)
