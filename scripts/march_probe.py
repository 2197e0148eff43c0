"""Kernel-level timing of the traversal kernels (march, offsets + expand) on the BASELINE configs, CUDA events,
L2 flushed between launches.  `python scripts/march_probe.py [reps]`; used with ncu for the per-kernel rows in profiles/."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfacc_b200 as nfa  # noqa: E402
from nerfacc_b200 import scenes  # noqa: E402
from nerfacc_b200.grid import _MarchJob  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    fn()
    tot = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot.append(a.elapsed_time(b) * 1e3)
    tot.sort()
    return {"median_us": round(tot[len(tot) // 2], 1), "min_us": round(tot[0], 1)}


for name, res, R, stratified in (("config2 128^3 65536 rays", 128, 65536, False),
                                 ("config2 stratified (per-ray near planes)", 128, 65536, True),
                                 ("config3 256^3 1048576 rays", 256, 1048576, False)):
    ro, rd = scenes.ball_rays(R)
    est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=res).to(dev)
    est.binaries = torch.from_numpy(scenes.ball_grid(res)).to(dev)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    ri, ts, te = est.sampling(ro, rd, render_step_size=scenes.BALL_STEP)
    N = ri.numel()
    near = far = None
    if stratified:
        near = (torch.rand(R, device=dev) * scenes.BALL_STEP).contiguous()
        far = torch.full((R,), 1e10, device=dev)
    job = _MarchJob(ro, rd, est.binaries, est.aabbs, near, far, scenes.BALL_STEP, None, None, None,
                    want_intervals=False, want_terminate=False, near_plane=0.0, far_plane=1e10)
    out = {"case": name, "n_samples": N, "march": timed(job._launch_march), "expand (+packed_info)": timed(lambda: job._expand_samples(N)),
           "sampling()": timed(lambda: est.sampling(ro, rd, render_step_size=scenes.BALL_STEP))}
    job.sc.busy = False
    print(json.dumps(out), flush=True)
    del ri, ts, te, job
