"""One launch of every nfa:: kernel on BASELINE-sized inputs, for `ncu --set full -k regex:nfa` captures:
config-2 step (march, offsets, expand, composite fwd/bwd), the visibility filter, the generic traversal,
standalone scans / pack_info, grid maintenance, config-4 importance sampling."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfacc_b200 as nfa  # noqa: E402
from nerfacc_b200 import scenes  # noqa: E402
from nerfacc_b200.data_specs import RayIntervals  # noqa: E402

dev = torch.device("cuda:0")
R = 65536
ro, rd = scenes.ball_rays(R)
est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
est.occs = est.binaries.float().flatten() * 0.5
tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
what = sys.argv[1] if len(sys.argv) > 1 else "all"

# -- config 2 step
ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
N = ri.numel()
sig = (5 * torch.rand(N, device=dev)).requires_grad_(True)
rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
col, op, dep, ex = nfa.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
torch.nn.functional.mse_loss(col, torch.rand(R, 3, device=dev)).backward()
if what == "all":
    # -- f1: visibility filter + compaction
    est.sampling(tro, trd, sigma_fn=lambda a, b, c: 3.0 + 0.0 * a, render_step_size=scenes.BALL_STEP, early_stop_eps=1e-2,
                 alpha_thre=1e-2)
    # -- f3: bounded test-mode marching (generic kernel), one round of the inference loop
    nfa.traverse_grids(tro, trd, est.binaries, est.aabbs, step_size=scenes.BALL_STEP, traverse_steps_limit=4,
                       over_allocate=True, rays_mask=torch.ones(R, dtype=torch.bool, device=dev))
    # -- standalone scans over 8.5 M elements, both addressings, and pack_info
    x = torch.rand(N, device=dev)
    pi = nfa.pack_info(ri, R)
    nfa.exclusive_sum(x, indices=ri)
    nfa.exclusive_sum(x, packed_info=pi)
    nfa.inclusive_prod(x, indices=ri)
    ri2 = ri.clone()  # no stashed packed_info: the real pack_info kernels run
    nfa.pack_info(ri2, R)
    nfa.accumulate_along_rays(x, None, ri2, R)
    # -- f4: grid maintenance
    est.train()
    est._update(step=1000, occ_eval_fn=lambda p: torch.exp(-6.0 * (p * p).sum(-1, keepdim=True)) * 0.05)
    # -- config 4: importance sampling 262144 x 64 -> 32
    Rp = 262144
    vals = torch.sort(torch.rand(Rp, 65, device=dev), dim=-1).values
    cdfs = torch.sort(torch.rand(Rp, 65, device=dev), dim=-1).values
    cdfs[:, 0], cdfs[:, -1] = 0.0, 1.0
    from nerfacc_b200.pdf import importance_sampling
    importance_sampling(RayIntervals(vals=vals), cdfs, 32, stratified=False)
    importance_sampling(RayIntervals(vals=vals), cdfs, 32, stratified=True)
torch.cuda.synchronize()
print("N", N)
