"""Run the UNMODIFIED reference CUDA build (baseline/_ref) on one named case and dump its outputs.

    python tests/ref_runner.py <case> <out.npz>

Used by tests/test_gpu_live_reference.py in a SUBPROCESS, so the reference's `nerfacc` package and this
repo's `nerfacc` alias never meet in one interpreter.  Nothing of nerfacc_b200 is imported here except the
numpy-only scene helpers (loaded by file path).  The same case definitions (`CASES`, `make_inputs`,
`field_sigma`, ...) are imported by the test to build identical inputs for the product.

Large per-sample arrays are stored raw when they fit comfortably (config 2: 8.5 M samples) and as SHA-256
digests + per-ray `packed_info` otherwise (config 3).
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("nfa_scenes", os.path.join(ROOT, "nerfacc_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(scenes)

CASES = {
    # BASELINE.json config 2: 128^3 ball grid, 65 536 rays, fwd + bwd
    "config2": dict(kind="train", res=128, levels=1, rays=65536, step=scenes.BALL_STEP),
    # config 3: 256^3 grid, inference only (262 144 of the 1 048 576 rays: same distribution, 1/4 of the time)
    "config3": dict(kind="infer", res=256, levels=1, rays=262144, step=scenes.BALL_STEP),
    # config 4: importance sampling 262 144 rays x 64 -> 32, both stratified modes
    "config4": dict(kind="pdf", rays=262144, n_in=64, n_out=32),
    # f1: sampling(sigma_fn / alpha_fn, early_stop_eps, alpha_thre) -- visibility filter
    "vis_ball": dict(kind="vis", res=128, levels=1, rays=16384, step=scenes.BALL_STEP, eps=1e-2, thre=1e-2, occ_mean=0.3),
    "vis_lvl4": dict(kind="vis", res=64, levels=4, rays=8192, step=1e-2, eps=1e-3, thre=5e-3, occ_mean=0.2),
    "vis_alpha": dict(kind="vis", res=128, levels=1, rays=16384, step=scenes.BALL_STEP, eps=1e-2, thre=2e-2, occ_mean=0.5,
                      alpha=True),
}


def sha(t) -> str:
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def grid_for(case) -> np.ndarray:
    """[levels, res, res, res] bool: a ball at level 0, random fragments inside bigger balls further out."""
    res, levels = case["res"], case["levels"]
    out = [scenes.ball_grid(res)[0]]
    rng = np.random.default_rng(7)
    for lv in range(1, levels):
        shell = scenes.ball_grid(res, radius=0.8)[0] & ~scenes.ball_grid(res, radius=0.3)[0]
        out.append(shell & (rng.random((res, res, res)) > 0.5))
    return np.stack(out)


def make_inputs(case):
    ro, rd = scenes.ball_rays(case["rays"], seed=42)
    return ro, rd, grid_for(case), scenes.nested_aabbs(case["levels"])


def field_sigma(t_starts, t_ends, ray_indices):
    """Deterministic stand-in for a density field: same torch ops on both sides => same bits."""
    tm = (t_starts + t_ends) * 0.5
    return 4.0 + 3.0 * torch.sin(7.0 * tm + 0.37 * (ray_indices % 1024).to(tm.dtype))


def field_alpha(t_starts, t_ends, ray_indices):
    tm = (t_starts + t_ends) * 0.5
    return 0.03 + 0.025 * torch.sin(5.0 * tm + 0.11 * (ray_indices % 512).to(tm.dtype))


def train_fields(n, n_rays, seed=43):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return 5 * torch.rand(n, generator=g), torch.rand(n, 3, generator=g), torch.rand(n_rays, 3, generator=g)


def pdf_inputs(case, seed=42):
    g = torch.Generator(device="cpu").manual_seed(seed)
    R, S = case["rays"], case["n_in"]
    vals = torch.sort(torch.rand(R, S + 1, generator=g), dim=-1).values
    cdfs = torch.sort(torch.rand(R, S + 1, generator=g), dim=-1).values
    cdfs[:, 0], cdfs[:, -1] = 0.0, 1.0
    return vals, cdfs


def run_case(nf, name, dev):
    """Run `name` through package `nf` (reference or product: same public API). Returns {key: np.ndarray | str}."""
    case = CASES[name]
    N = lambda t: t.detach().cpu().numpy()
    out = {}
    if case["kind"] == "pdf":
        from importlib import import_module
        RayIntervals = import_module(nf.__name__ + ".data_specs").RayIntervals
        importance_sampling = import_module(nf.__name__ + ".pdf").importance_sampling
        vals, cdfs = pdf_inputs(case)
        vals, cdfs = vals.to(dev), cdfs.to(dev)
        for strat in (False, True):
            torch.manual_seed(1234)
            iv, sm = importance_sampling(RayIntervals(vals=vals), cdfs, case["n_out"], stratified=strat)
            out[f"iv_{int(strat)}"] = N(iv.vals)
            out[f"sm_{int(strat)}"] = N(sm.vals)
        return out
    ro, rd, bins, aabbs = make_inputs(case)
    est = nf.OccGridEstimator(torch.from_numpy(aabbs[0]), resolution=case["res"], levels=case["levels"]).to(dev)
    est.binaries = torch.from_numpy(bins).to(dev)
    tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    R = case["rays"]
    if case["kind"] == "vis":
        est.occs.fill_(case["occ_mean"])
        kw = dict(alpha_fn=field_alpha) if case.get("alpha") else dict(sigma_fn=field_sigma)
        ri0, ts0, te0 = est.sampling(tro, trd, render_step_size=case["step"])
        ri, ts, te = est.sampling(tro, trd, render_step_size=case["step"], early_stop_eps=case["eps"],
                                  alpha_thre=case["thre"], **kw)
        out.update(n_before=np.int64(ri0.numel()), ri=N(ri).astype(np.int32), ts=N(ts), te=N(te))
        return out
    ri, ts, te = est.sampling(tro, trd, render_step_size=case["step"])
    n = ri.numel()
    pi = nf.pack_info(ri, R)
    out.update(n=np.int64(n), packed_info=N(pi))
    if case["kind"] == "train":
        sig, rgb, tgt = train_fields(n, R)
        sig, rgb, tgt = sig.to(dev).requires_grad_(True), rgb.to(dev).requires_grad_(True), tgt.to(dev)
        col, op, dep, ex = nf.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
        torch.nn.functional.mse_loss(col, tgt).backward()
        out.update(ri=N(ri).astype(np.int32), ts=N(ts), te=N(te), colors=N(col), opacities=N(op), depths=N(dep),
                   weights=N(ex["weights"]), trans=N(ex["trans"]), alphas=N(ex["alphas"]), g_sig=N(sig.grad),
                   g_rgb=N(rgb.grad))
    else:
        with torch.no_grad():
            sig, rgb, _ = train_fields(n, R)
            sig, rgb = sig.to(dev), rgb.to(dev)
            col, op, dep, ex = nf.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
        out.update(sha_ri=sha(ri), sha_ts=sha(ts), sha_te=sha(te), colors=N(col), opacities=N(op), depths=N(dep))
    return out


if __name__ == "__main__":
    name, path = sys.argv[1], sys.argv[2]
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import nerfacc as ref
    assert "baseline/_ref" in ref.__file__.replace("\\", "/"), ref.__file__
    res = run_case(ref, name, torch.device("cuda:0"))
    np.savez(path, **{k: np.asarray(v) for k, v in res.items()})
    print("ok", name, {k: getattr(v, "shape", v) for k, v in res.items()})
