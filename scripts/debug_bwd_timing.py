import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfacc_b200 as nfa
from nerfacc_b200 import scenes
dev = torch.device("cuda:0")
R = 65536
ro, rd = scenes.ball_rays(R)
est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
ri, ts, te = est.sampling(torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev), render_step_size=scenes.BALL_STEP)
N = ri.numel()
sig = (5 * torch.rand(N, device=dev)).requires_grad_(True); rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
gcol = torch.rand(R, 3, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for mode in ["backward(gcol)", "autograd.grad", "sum-loss"]:
    for rep in range(4):
        sig.grad = None; rgb.grad = None
        col = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))[0]
        flush.fill_(1)
        torch.cuda.synchronize()
        a, c = ev(), ev()
        t0 = time.perf_counter()
        a.record()
        if mode == "backward(gcol)": col.backward(gcol)
        elif mode == "autograd.grad": torch.autograd.grad(col, [sig, rgb], gcol)
        else: (col * gcol).sum().backward()
        c.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(mode, rep, f"gpu {a.elapsed_time(c)*1e3:.1f} us  host {1e6*(t1-t0):.1f} us", flush=True)
