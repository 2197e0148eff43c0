"""One sampling + rendering fwd/bwd step for ncu captures (config 2 workload)."""
import sys, os
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
tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    N = ri.numel()
    sig = (5 * torch.rand(N, device=dev)).requires_grad_(True)
    rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
    col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
    torch.nn.functional.mse_loss(col, torch.rand(R, 3, device=dev)).backward()
torch.cuda.synchronize()
print("N", N)
