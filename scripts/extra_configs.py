"""BASELINE.json configs 3 and 4 on one GPU, for either implementation (same public API):

    python scripts/extra_configs.py ours            # this repo (nerfacc_b200)
    python scripts/extra_configs.py reference-cuda  # the unmodified reference CUDA build in baseline/_ref

config 3: 256^3 occ-grid, 1 048 576 rays, inference-only (sampling + no-grad compositing)
config 4: importance sampling, 262 144 rays x 64 -> 32 samples (kernel alone, and one PropNet level end to end)
Times with CUDA events after warm-up; prints one JSON line per config and appends them to
gpurun_out/extra_configs_<impl>.json.  Informational (bench.py stays on config 2).
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
if impl == "reference-cuda":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import nerfacc as nf
    assert "baseline/_ref" in nf.__file__
    from nerfacc.data_specs import RayIntervals
    from nerfacc.estimators.prop_net import PropNetEstimator
    from nerfacc.pdf import importance_sampling
    from nerfacc.volrend import accumulate_along_rays_
else:
    sys.path.insert(0, ROOT)
    import nerfacc_b200 as nf
    from nerfacc_b200.data_specs import RayIntervals
    from nerfacc_b200.estimators.prop_net import PropNetEstimator
    from nerfacc_b200.pdf import importance_sampling
    from nerfacc_b200.volrend import accumulate_along_rays_
_spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "nerfacc_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(scenes)

dev = torch.device("cuda:0")
lines = []
ONLY = [x for x in os.environ.get("NFA_EXTRA_ONLY", "").split(",") if x]


def want(tag):
    return not ONLY or tag in ONLY



def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def emit(**kw):
    kw["impl"] = impl
    lines.append(kw)
    print(json.dumps(kw), flush=True)


# ---------------------------------------------------------------- config 2 + visibility filter (row f1)
R2 = 65536
ro2, rd2 = scenes.ball_rays(R2)
est2 = nf.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est2.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
est2.occs = est2.binaries.float().flatten() * 0.5
tro2, trd2 = torch.from_numpy(ro2).to(dev), torch.from_numpy(rd2).to(dev)


def sigma_fn(t_starts, t_ends, ray_indices):
    return 3.0 + 0.0 * t_starts  # constant density: rays saturate after ~3 units of optical depth


with torch.no_grad():
    ri2, _, _ = est2.sampling(tro2, trd2, render_step_size=scenes.BALL_STEP)
    rv2, _, _ = est2.sampling(tro2, trd2, sigma_fn=sigma_fn, render_step_size=scenes.BALL_STEP, early_stop_eps=1e-2,
                              alpha_thre=1e-2)
    t_v = timed(lambda: est2.sampling(tro2, trd2, sigma_fn=sigma_fn, render_step_size=scenes.BALL_STEP,
                                      early_stop_eps=1e-2, alpha_thre=1e-2), 20)
    t_p = timed(lambda: est2.sampling(tro2, trd2, render_step_size=scenes.BALL_STEP), 20)
emit(config="2 + f1: sampling with sigma_fn visibility filter (early_stop_eps=1e-2, alpha_thre=1e-2), 65536 rays",
     n_before=ri2.numel(), n_kept=rv2.numel(), filtered_us=t_v * 1e6, plain_us=t_p * 1e6)
del est2, ri2, rv2
torch.cuda.empty_cache()

if want('c3'):
    # ---------------------------------------------------------------- config 3
    R3, G3 = 1 << 20, 256
    ro, rd = scenes.ball_rays(R3, seed=7)
    est = nf.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=G3).to(dev)
    est.binaries = torch.from_numpy(scenes.ball_grid(G3)).to(dev)
    tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    step3 = 2.6e-3  # half of config 2's step on a grid twice as fine
    with torch.no_grad():
        ri, ts, te = est.sampling(tro, trd, render_step_size=step3)
        N3 = ri.numel()
        sig = 5 * torch.rand(N3, device=dev)
        rgb = torch.rand(N3, 3, device=dev)

        def infer():
            ri_, ts_, te_ = est.sampling(tro, trd, render_step_size=step3)
            return nf.rendering(ts_, te_, ri_, n_rays=R3, rgb_sigma_fn=lambda a, b, c: (rgb, sig))

        def samp():
            return est.sampling(tro, trd, render_step_size=step3)

        def comp():
            return nf.rendering(ts, te, ri, n_rays=R3, rgb_sigma_fn=lambda a, b, c: (rgb, sig))

        t_all, t_s, t_c = timed(infer, 10), timed(samp, 10), timed(comp, 10)
    emit(config="3: 256^3 occ-grid, 1048576 rays, inference-only", n_samples=N3, samples_per_ray=N3 / R3,
         step_ms=t_all * 1e3, sampling_ms=t_s * 1e3, compositing_ms=t_c * 1e3, gsamples_per_s=N3 / t_all / 1e9)
    del ri, ts, te, sig, rgb
    torch.cuda.empty_cache()

if want('c4'):
    # ---------------------------------------------------------------- config 4
    R4 = 262144
    torch.manual_seed(3)
    edges = torch.sort(torch.rand(R4, 65, device=dev), -1)[0]
    edges[:, 0], edges[:, -1] = 0.0, 1.0
    w = torch.rand(R4, 64, device=dev) ** 4 + 1e-3
    cdfs = torch.cat([torch.zeros(R4, 1, device=dev), torch.cumsum(w, -1)], -1)
    cdfs = (cdfs / cdfs[:, -1:]).contiguous()
    iv_in = RayIntervals(vals=edges)
    for strat in (False, True):
        t = timed(lambda: importance_sampling(iv_in, cdfs, 32, strat), 50, 5)
        # algorithmic bytes: read edges + cdfs (2 x 65 floats), write 32 centres + 33 edges, per ray
        by = R4 * 4 * (2 * 65 + 32 + 33)
        emit(config="4: importance_sampling 262144 rays x 64 -> 32", stratified=strat, us=t * 1e6,
             gsamples_per_s=R4 * 32 / t / 1e9, algorithmic_gb_per_s=by / t / 1e9)


    def prop_fn(t_starts, t_ends):
        mid = (t_starts + t_ends) * 0.5
        return 4.0 * torch.exp(-((mid - 3.0) / 1.0) ** 2)


    pn = PropNetEstimator().to(dev)
    t = timed(lambda: pn.sampling([prop_fn], [64], 32, n_rays=R4, near_plane=0.2, far_plane=50.0, sampling_type="lindisp",
                                  stratified=True), 20, 3)
    emit(config="4: PropNetEstimator.sampling 262144 rays, one proposal level 64 -> 32 final", ms=t * 1e3,
         gsamples_per_s=R4 * 32 / t / 1e9)

# ---------------------------------------------------------------- standalone scans at config-2 size (K5: CUB by-key vs ours)
est2 = nf.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est2.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
with torch.no_grad():
    ri2, ts2, te2 = est2.sampling(tro2, trd2, render_step_size=scenes.BALL_STEP)
N2 = ri2.numel()
x2 = torch.rand(N2, device=dev)
ri_plain = ri2.clone()  # carries no stashed segments: the key-addressed kernels run
pi2 = nf.pack_info(ri_plain, R2)
emit(config="scans: N = 8.5 M samples in 65536 segments", n=N2,
     exclusive_sum_indices_us=timed(lambda: nf.exclusive_sum(x2, indices=ri_plain), 30, 5) * 1e6,
     inclusive_sum_indices_us=timed(lambda: nf.inclusive_sum(x2, indices=ri_plain), 30, 5) * 1e6,
     exclusive_prod_indices_us=timed(lambda: nf.exclusive_prod(x2, indices=ri_plain), 30, 5) * 1e6,
     exclusive_sum_packed_us=timed(lambda: nf.exclusive_sum(x2, packed_info=pi2), 30, 5) * 1e6,
     pack_info_us=timed(lambda: nf.pack_info(ri_plain, R2), 30, 5) * 1e6,
     note="8 B/sample algorithmic (read + write f32) + 8 B/sample of int64 keys on the `indices` route")

# ---------------------------------------------------------------- f3: test-mode marching with early termination
# the loop of /root/reference/examples/utils.py:267-439 (render_image_with_occgrid_test) on config-2 rays, through
# the public API only: traverse_grids(limit, over_allocate, rays_mask) -> render_weight_from_density(prefix_trans)
# -> 3 x accumulate_along_rays_
def test_mode_render(estimator, rays_o, rays_d, step, sigma_const=20.0, early_stop_eps=1e-4, max_samples=1024):
    n = rays_o.shape[0]
    opacity = torch.zeros(n, 1, device=dev)
    depth = torch.zeros(n, 1, device=dev)
    rgb = torch.zeros(n, 3, device=dev)
    ray_mask = torch.ones(n, device=dev).bool()
    near_planes = torch.zeros(n, device=dev)
    far_planes = torch.full((n,), 1e10, device=dev)
    t_mins, t_maxs, hits = nf.ray_aabb_intersect(rays_o, rays_d, estimator.aabbs)
    t_sorted = torch.cat([t_mins, t_maxs], -1)
    t_indices = torch.arange(0, 2, device=dev, dtype=torch.int64).expand(n, 2)
    opc_thre = 1 - early_stop_eps
    iter_samples = total = rounds = 0
    while iter_samples < max_samples:
        n_alive = int(ray_mask.sum().item())
        if n_alive == 0:
            break
        n_samples = max(min(n // n_alive, 64), 1)
        iter_samples += n_samples
        intervals, samples, term = nf.traverse_grids(rays_o, rays_d, estimator.binaries, estimator.aabbs, near_planes,
                                                     far_planes, step, 0.0, n_samples, True, ray_mask, t_sorted,
                                                     t_indices, hits)
        t_starts = intervals.vals[intervals.is_left]
        t_ends = intervals.vals[intervals.is_right]
        ray_indices = samples.ray_indices[samples.is_valid]
        sigmas = torch.full_like(t_starts, sigma_const)
        rgbs = torch.sigmoid(t_starts)[:, None].expand(-1, 3)
        weights, _, _ = nf.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n,
                                                      prefix_trans=1 - opacity[ray_indices].squeeze(-1))
        accumulate_along_rays_(weights, values=rgbs, ray_indices=ray_indices, outputs=rgb)
        accumulate_along_rays_(weights, values=None, ray_indices=ray_indices, outputs=opacity)
        accumulate_along_rays_(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices,
                                  outputs=depth)
        near_planes = term
        ray_mask = torch.logical_and(opacity.view(-1) <= opc_thre, samples.packed_info[:, 1] == n_samples)
        total += ray_indices.shape[0]
        rounds += 1
    return rgb, opacity, total, rounds


with torch.no_grad():
    rgb_t, op_t, total_t, rounds_t = test_mode_render(est2, tro2, trd2, scenes.BALL_STEP)
    t_tm = timed(lambda: test_mode_render(est2, tro2, trd2, scenes.BALL_STEP), 5, 2)
    one = timed(lambda: nf.traverse_grids(tro2, trd2, est2.binaries, est2.aabbs, torch.zeros(R2, device=dev),
                                          torch.full((R2,), 1e10, device=dev), scenes.BALL_STEP, 0.0, 4, True,
                                          torch.ones(R2, dtype=torch.bool, device=dev)), 20, 3)
emit(config="f3: test-mode rendering loop (bounded marching + prefix_trans + in-place accumulate), 65536 rays, sigma = 20",
     total_samples=total_t, rounds=rounds_t, loop_ms=t_tm * 1e3, checksum_rgb=float(rgb_t.double().sum()),
     checksum_opacity=float(op_t.double().sum()), one_bounded_traverse_call_us=one * 1e6)

# ---------------------------------------------------------------- f4: grid maintenance (128^3, one level)
est4 = nf.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est4.train()


def occ_eval_fn(x):
    return torch.exp(-6.0 * (x * x).sum(-1, keepdim=True)) * 0.05


est4._update(step=0, occ_eval_fn=occ_eval_fn)
t_warm = timed(lambda: est4._update(step=0, occ_eval_fn=occ_eval_fn), 10, 2)
t_samp = timed(lambda: est4._update(step=1000, occ_eval_fn=occ_eval_fn), 10, 2)
t_next = timed(lambda: (est4._update(step=1000, occ_eval_fn=occ_eval_fn),
                        est4.sampling(tro2, trd2, render_step_size=scenes.BALL_STEP)), 10, 2)
emit(config="f4: OccGridEstimator._update, 128^3", all_cells_us=t_warm * 1e6, sampled_cells_us=t_samp * 1e6,
     update_then_sampling_us=t_next * 1e6, occupied=int(est4.binaries.sum()))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"extra_configs_{impl}.json"), "w") as f:
    for ln in lines:
        f.write(json.dumps(ln) + "\n")
