"""Time the UNMODIFIED reference CUDA build (baseline/_ref) on the config-2 workload.
Informational: the "reference's own CUDA build on 1 GPU" bar of BASELINE.json."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import nerfacc as ref
import importlib.util
_spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "nerfacc_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(scenes)
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = 20
ro, rd = scenes.ball_rays(R)
est = ref.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
tro, trd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for _ in range(3): ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
N = ri.numel()
torch.cuda.synchronize(); e0 = ev()
for _ in range(iters): ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
e1 = ev(); torch.cuda.synchronize(); t_samp = e0.elapsed_time(e1) / iters
sig = (5 * torch.rand(N, device=dev)).requires_grad_(True); rgb = torch.rand(N, 3, device=dev).requires_grad_(True)
tgt = torch.rand(R, 3, device=dev)
def step():
    ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
    col, op, dep, ex = ref.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
    sig.grad = None; rgb.grad = None
    torch.nn.functional.mse_loss(col, tgt).backward()
for _ in range(3): step()
torch.cuda.synchronize(); e0 = ev()
for _ in range(iters): step()
e1 = ev(); torch.cuda.synchronize(); t_step = e0.elapsed_time(e1) / iters
ri, ts, te = est.sampling(tro, trd, render_step_size=scenes.BALL_STEP)
torch.cuda.synchronize(); e0 = ev()
for _ in range(iters):
    col, op, dep, ex = ref.rendering(ts, te, ri, n_rays=R, rgb_sigma_fn=lambda a, b, c: (rgb, sig))
    sig.grad = None; rgb.grad = None
    torch.nn.functional.mse_loss(col, tgt).backward()
e1 = ev(); torch.cuda.synchronize(); t_rend = e0.elapsed_time(e1) / iters
out = dict(impl="reference-cuda", n_rays=R, n_samples=N, sampling_us=t_samp * 1e3, render_fwd_bwd_us=t_rend * 1e3,
           step_us=t_step * 1e3, gsamples_per_s=N / t_step / 1e6)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "ref_cuda_timing.json"), "w").write(json.dumps(out) + "\n")
