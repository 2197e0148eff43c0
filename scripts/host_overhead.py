"""Host-side cost of each phase of a step (GPU synchronised between phases, so only launch/Python time is seen)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfacc_b200 as nfa
from nerfacc_b200 import scenes, _lib
dev = torch.device("cuda:0")
R = 65536
ro, rd = scenes.ball_rays(R)
est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
N = ri.numel()
sig = (5 * torch.rand(N, device=dev)).requires_grad_(True); rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
tgt = torch.rand(R, 3, device=dev)
field = lambda a, b, c: (rgb, sig)
T = {}
def lap(name, t0):
    torch.cuda.synchronize()
    T.setdefault(name, []).append(time.perf_counter() - t0)
for it in range(60):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP); t1 = time.perf_counter()
    T.setdefault("sampling(host-return)", []).append(t1 - t0); torch.cuda.synchronize()
    t0 = time.perf_counter(); col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=field); t1 = time.perf_counter()
    T.setdefault("rendering(host-return)", []).append(t1 - t0); torch.cuda.synchronize()
    t0 = time.perf_counter(); loss = torch.nn.functional.mse_loss(col, tgt); t1 = time.perf_counter()
    T.setdefault("mse(host-return)", []).append(t1 - t0); torch.cuda.synchronize()
    sig.grad = None; rgb.grad = None
    t0 = time.perf_counter()
    with torch.autograd.set_multithreading_enabled(False):
        loss.backward()
    t1 = time.perf_counter()
    T.setdefault("backward(host-return)", []).append(t1 - t0); torch.cuda.synchronize()
for k, v in T.items():
    v = np.array(v[10:]) * 1e6
    print(f"{k:28s} median {np.median(v):7.1f} us   min {v.min():7.1f}")
# finer: pieces of rendering
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=field)
    loss = torch.nn.functional.mse_loss(col, tgt)
    sig.grad = None; rgb.grad = None
    with torch.autograd.set_multithreading_enabled(False):
        loss.backward()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(30)
st.sort_stats("tottime").print_stats(30)
