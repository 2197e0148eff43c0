"""Ad-hoc first GPU run: parity of the native path vs the oracle + first timings.
Run from the repo root on a GPU box; writes gpurun_out/first_contact.log."""
import os, sys, time, traceback
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "first_contact.log"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n"); LOG.flush()


import nerfacc_b200 as nfa
from nerfacc_b200 import scenes
from oracle import oracle as orc

dev = torch.device("cuda:0")
log(torch.cuda.get_device_name(0), "oracle threads", orc.num_threads())


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def check_sampling(name, ro, rd, bins, aabbs, **kw):
    est = nfa.OccGridEstimator(torch.from_numpy(aabbs[0]), resolution=list(bins.shape[1:]), levels=bins.shape[0]).to(dev)
    est.binaries = T(bins)
    ri, ts, te = est.sampling(T(ro), T(rd), **kw)
    torch.cuda.synchronize()
    o_ri, o_ts, o_te, o_pi = orc.occgrid_sampling(ro, rd, bins, aabbs, near_plane=kw.get("near_plane", 0.0),
                                                  far_plane=kw.get("far_plane", 1e10),
                                                  t_min=None if kw.get("t_min") is None else kw["t_min"].cpu().numpy(),
                                                  t_max=None if kw.get("t_max") is None else kw["t_max"].cpu().numpy(),
                                                  render_step_size=kw.get("render_step_size", 1e-3))
    ok = (np.array_equal(ri.cpu().numpy(), o_ri) and np.array_equal(ts.cpu().numpy(), o_ts)
          and np.array_equal(te.cpu().numpy(), o_te))
    pi = ri._nfa_packed[0].cpu().numpy()
    ok_pi = np.array_equal(pi, o_pi)
    log(f"[sampling] {name}: N={len(o_ri)} mine={ri.numel()} exact={ok} packed_info={ok_pi}")
    return ok and ok_pi


results = {}
try:
    R = 4096
    ro, rd = scenes.ball_rays(R)
    bins = scenes.ball_grid(128)
    aabbs = scenes.nested_aabbs(1)
    results["ball"] = check_sampling("ball128 R=4096", ro, rd, bins, aabbs, render_step_size=scenes.BALL_STEP)
    results["ball_again"] = check_sampling("ball128 R=4096 (hinted)", ro, rd, bins, aabbs, render_step_size=scenes.BALL_STEP)
    rng = np.random.default_rng(7)
    frag = bins & (rng.random(bins.shape) > 0.5)
    results["frag"] = check_sampling("fragmented ball", ro, rd, frag, aabbs, render_step_size=scenes.BALL_STEP)
    R2 = 300
    ro2 = rng.standard_normal((R2, 3)).astype(np.float32)
    rd2 = rng.standard_normal((R2, 3)).astype(np.float32); rd2 /= np.linalg.norm(rd2, axis=1, keepdims=True)
    bins4 = rng.random((4, 32, 32, 32)) > 0.5
    results["lvl4"] = check_sampling("4-level random 32^3", ro2, rd2, bins4, scenes.nested_aabbs(4), render_step_size=1e-2)
    tmin = torch.rand(R2, device=dev); tmax = tmin + torch.rand(R2, device=dev)
    results["tminmax"] = check_sampling("4-level t_min/t_max", ro2, rd2, bins4, scenes.nested_aabbs(4), render_step_size=1e-2,
                                        near_plane=0.15, far_plane=0.85 * 4, t_min=tmin, t_max=tmax)
    bo = rng.random((2, 30, 17, 5)) > 0.3
    results["odd"] = check_sampling("2-level 30x17x5", ro2, rd2, bo, scenes.nested_aabbs(2), render_step_size=4e-3)
except Exception:
    log("SAMPLING EXCEPTION\n" + traceback.format_exc())

# traverse_grids (intervals) vs oracle
try:
    iv, sm, term = nfa.traverse_grids(T(ro2), T(rd2), T(bins4), T(scenes.nested_aabbs(4)), step_size=1e-2)
    o_iv, o_sm, o_term = orc.traverse_grids(ro2, rd2, bins4, scenes.nested_aabbs(4), step_size=1e-2)
    ok = all([
        np.array_equal(iv.vals.cpu().numpy(), o_iv["vals"]), np.array_equal(iv.ray_indices.cpu().numpy(), o_iv["ray_indices"]),
        np.array_equal(iv.is_left.cpu().numpy(), o_iv["is_left"]), np.array_equal(iv.is_right.cpu().numpy(), o_iv["is_right"]),
        np.array_equal(iv.packed_info.cpu().numpy(), o_iv["packed_info"]),
        np.array_equal(sm.vals.cpu().numpy(), o_sm["vals"]), np.array_equal(sm.ray_indices.cpu().numpy(), o_sm["ray_indices"]),
        np.array_equal(sm.packed_info.cpu().numpy(), o_sm["packed_info"]), bool(sm.is_valid.all()),
    ])
    m = ~np.isnan(o_term)
    ok_t = np.array_equal(term.cpu().numpy()[m], o_term[m])
    log(f"[traverse_grids] intervals/samples exact={ok} terminate={ok_t} E={len(o_iv['vals'])}")
    results["traverse"] = ok and ok_t
except Exception:
    log("TRAVERSE EXCEPTION\n" + traceback.format_exc())

# compositing vs oracle
try:
    R = 4096
    ro, rd = scenes.ball_rays(R)
    est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
    est.binaries = T(scenes.ball_grid(128))
    ri, ts, te = est.sampling(T(ro), T(rd), render_step_size=scenes.BALL_STEP)
    N = ri.numel()
    g = torch.Generator(device="cpu").manual_seed(43)
    sig = (5 * torch.rand(N, generator=g)).to(dev).requires_grad_(True)
    rgb = torch.rand(N, 3, generator=g).to(dev).requires_grad_(True)
    bk = torch.tensor([0.2, 0.5, 0.9], device=dev)
    col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig), render_bkgd=bk)
    gC = torch.rand(R, 3, generator=g).to(dev); gO = torch.rand(R, 1, generator=g).to(dev); gD = torch.rand(R, 1, generator=g).to(dev)
    loss = (col * gC).sum() + (op * gO).sum() + (dep * gD).sum()
    loss.backward()
    pi = ri._nfa_packed[0].cpu().numpy()
    o = orc.composite(ts.cpu().numpy(), te.cpu().numpy(), sig.detach().cpu().numpy(), rgb.detach().cpu().numpy(), packed_info=pi,
                      render_bkgd=bk.cpu().numpy())
    gs, gr = orc.composite_backward(ts.cpu().numpy(), te.cpu().numpy(), sig.detach().cpu().numpy(), rgb.detach().cpu().numpy(), pi,
                                    gC=gC.cpu().numpy(), gO=gO.cpu().numpy().ravel(), gD=gD.cpu().numpy().ravel(), render_bkgd=bk.cpu().numpy())
    def md(a, b): return float(np.abs(a.detach().cpu().numpy().astype(np.float64) - b).max())
    errs = dict(weights=md(ex["weights"], o["weights"]), trans=md(ex["trans"], o["trans"]), alphas=md(ex["alphas"], o["alphas"]),
                colors=md(col, o["colors"]), opac=md(op, o["opacities"]), depth=md(dep, o["depths"]),
                g_sigma=md(sig.grad, gs), g_rgb=md(rgb.grad, gr))
    log("[composite] max abs err vs oracle:", errs)
    results["composite"] = all(v < 1e-5 for k, v in errs.items() if not k.startswith("g_")) and errs["g_sigma"] < 1e-4 and errs["g_rgb"] < 1e-5
    # general backward (grads on weights / trans / alphas)
    sig.grad = None; rgb.grad = None
    w, Tt, a = nfa.render_weight_from_density(ts, te, sig, ray_indices=ri, n_rays=R)
    gW = torch.rand(N, generator=g).to(dev); gT = torch.rand(N, generator=g).to(dev); gA = torch.rand(N, generator=g).to(dev)
    ((w * gW).sum() + (Tt * gT).sum() + (a * gA).sum()).backward()
    gs2, _ = orc.composite_backward(ts.cpu().numpy(), te.cpu().numpy(), sig.detach().cpu().numpy(), None, pi,
                                    gW=gW.cpu().numpy(), gT=gT.cpu().numpy(), gA=gA.cpu().numpy())
    e2 = md(sig.grad, gs2)
    log("[composite] general bwd max abs err:", e2, "max|g|", float(np.abs(gs2).max()))
    results["composite_general"] = e2 < 1e-4
except Exception:
    log("COMPOSITE EXCEPTION\n" + traceback.format_exc())

# scans + pack_info vs oracle
try:
    g = torch.Generator(device="cpu").manual_seed(5)
    cnts = torch.randint(0, 300, (3000,), generator=g)
    starts = torch.cumsum(cnts, 0) - cnts
    pinfo = torch.stack([starts, cnts], -1).to(dev)
    n = int(cnts.sum())
    idx = torch.repeat_interleave(torch.arange(3000), cnts).to(dev)
    x = (torch.rand(n, generator=g) * 0.2 + 0.9).to(dev)
    ok = True
    for name, fn, ofn in [("inclusive_sum", nfa.inclusive_sum, orc.inclusive_sum), ("exclusive_sum", nfa.exclusive_sum, orc.exclusive_sum),
                          ("inclusive_prod", nfa.inclusive_prod, orc.inclusive_prod), ("exclusive_prod", nfa.exclusive_prod, orc.exclusive_prod)]:
        a = fn(x, packed_info=pinfo).cpu().numpy(); b = fn(x, indices=idx).cpu().numpy()
        o = ofn(x.cpu().numpy(), packed_info=pinfo.cpu().numpy())
        e1 = np.abs(a - o).max() / max(1, np.abs(o).max()); e2 = np.abs(b - o).max() / max(1, np.abs(o).max())
        log(f"[scan] {name}: rel err packed={e1:.2e} bykey={e2:.2e}")
        ok &= e1 < 1e-5 and e2 < 1e-5
    pk = nfa.pack_info(idx, 3000).cpu().numpy()
    ok_pk = np.array_equal(pk, pinfo.cpu().numpy())
    log("[pack_info] exact:", ok_pk)
    results["scan"] = ok and ok_pk
except Exception:
    log("SCAN EXCEPTION\n" + traceback.format_exc())

log("RESULTS", results)

# ---- timings (config 2) ----
try:
    R = 65536
    ro, rd = scenes.ball_rays(R)
    est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
    est.binaries = T(scenes.ball_grid(128))
    tro, trd = T(ro), T(rd)
    def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
    for _ in range(3): ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    N = ri.numel(); log("config2 N =", N, "spp", N / R)
    torch.cuda.synchronize(); t0 = time.time(); e0 = ev()
    for _ in range(20): ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    e1 = ev(); torch.cuda.synchronize(); t1 = time.time()
    log(f"sampling: {e0.elapsed_time(e1)/20*1e3:.1f} us/iter (events), wall {(t1-t0)/20*1e6:.1f} us")
    sig = (5 * torch.rand(N, device=dev)).requires_grad_(True); rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
    tgt = torch.rand(R, 3, device=dev)
    def step():
        ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
        col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
        loss = torch.nn.functional.mse_loss(col, tgt)
        sig.grad = None; rgb.grad = None
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.time(); e0 = ev()
    for _ in range(20): step()
    e1 = ev(); torch.cuda.synchronize(); t1 = time.time()
    ms = e0.elapsed_time(e1) / 20
    log(f"full step: {ms*1e3:.1f} us/iter (events), wall {(t1-t0)/20*1e6:.1f} us -> {N/ms/1e6:.2f} G samples/s")
    # per-stage
    ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    torch.cuda.synchronize(); e0 = ev()
    for _ in range(20): col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
    e1 = ev(); torch.cuda.synchronize(); log(f"rendering fwd: {e0.elapsed_time(e1)/20*1e3:.1f} us")
    e0 = ev()
    for _ in range(20):
        col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
        sig.grad = None; rgb.grad = None
        torch.nn.functional.mse_loss(col, tgt).backward()
    e1 = ev(); torch.cuda.synchronize(); log(f"rendering fwd+bwd: {e0.elapsed_time(e1)/20*1e3:.1f} us")
except Exception:
    log("TIMING EXCEPTION\n" + traceback.format_exc())
LOG.close()
