"""Per-warp time stamps of march_kernel on config 2 (nfa_debug_set_march_trace): when does each tile start, how long
does each warp take for staging + sort, for its rays, and until the tile's last barrier.  Measurement aid."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfacc_b200 as nfa  # noqa: E402
from nerfacc_b200 import _lib, scenes  # noqa: E402
from nerfacc_b200.grid import _MarchJob  # noqa: E402

dev = torch.device("cuda:0")
R = 65536
ro, rd = scenes.ball_rays(R)
est = nfa.OccGridEstimator(torch.from_numpy(scenes.ROI_AABB), resolution=128).to(dev)
est.binaries = torch.from_numpy(scenes.ball_grid(128)).to(dev)
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
est.sampling(ro, rd, render_step_size=scenes.BALL_STEP)
job = _MarchJob(ro, rd, est.binaries, est.aabbs, None, None, scenes.BALL_STEP, None, None, None, want_intervals=False,
                want_terminate=False, near_plane=0.0, far_plane=1e10)
lib = _lib.load()
n_tiles_max = 4096
trace = torch.zeros(n_tiles_max * 16 * 8, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    job._launch_march()
torch.cuda.synchronize()
lib.nfa_debug_set_march_trace(trace.data_ptr())
flush.fill_(1)
job._launch_march()
torch.cuda.synchronize()
lib.nfa_debug_set_march_trace(None)
t = trace.cpu().numpy().reshape(n_tiles_max, 16, 8)
used = t[:, :, 7] == 1
tiles = np.where(used.any(1))[0]
print("tiles", len(tiles), "warps per tile", int(used[tiles[0]].sum()))
t0 = t[tiles][:, :, 0].astype(np.float64)
t1 = t[tiles][:, :, 1].astype(np.float64)
w = used[tiles]
g0 = t0[w].min()
start = np.array([t0[i][w[i]].min() for i in range(len(tiles))]) - g0
end = np.array([t1[i][w[i]].max() for i in range(len(tiles))]) - g0
print("tile start (us after the first): min %.1f median %.1f max %.1f" % (start.min() / 1e3, np.median(start) / 1e3, start.max() / 1e3))
print("tile end   (us after the first start): min %.1f median %.1f max %.1f" % (end.min() / 1e3, np.median(end) / 1e3, end.max() / 1e3))
dur = (end - start) / 1e3
print("tile duration us: min %.1f median %.1f p90 %.1f max %.1f" % (dur.min(), np.median(dur), np.percentile(dur, 90), dur.max()))
clk = 1.965e3  # cycles per us
setup = t[tiles][:, :, 2][w] / clk
rays_done = t[tiles][:, :, 3][w] / clk
barrier = t[tiles][:, :, 4][w] / clk
print("per warp, us from its start: staging+sort median %.1f max %.1f | rays done median %.1f p90 %.1f max %.1f | past last barrier median %.1f max %.1f"
      % (np.median(setup), setup.max(), np.median(rays_done), np.percentile(rays_done, 90), rays_done.max(), np.median(barrier), barrier.max()))
walk1 = (t[tiles][:, :, 6] & 0xffffffffffff) / clk
loops = t[tiles][:, :, 6] >> 48
print("first walk pass done, by warp index (median us):", [round(float(np.median(walk1[:, k][w[:, k]])), 1) if w[:, k].any() else None for k in range(16)])
print("walk/lattice rounds, by warp index (median, max):", [(int(np.median(loops[:, k][w[:, k]])), int(loops[:, k][w[:, k]].max())) if w[:, k].any() else None for k in range(16)])
# per tile: which warp finishes last and when
rd_t = t[tiles][:, :, 3] / clk
last_warp = np.array([np.argmax(np.where(w[i], rd_t[i], -1)) for i in range(len(tiles))])
print("slowest warp of a tile (index histogram):", np.bincount(last_warp, minlength=16).tolist())
print("rays-done time by warp index (median over tiles):", [round(float(np.median(rd_t[:, k][w[:, k]])), 1) if w[:, k].any() else None for k in range(16)])
worst = np.argsort(-dur)[:5]
for i in worst:
    print("slow tile", int(tiles[i]), "smid", int(t[tiles[i], 0, 5]), "start %.1f dur %.1f" % (start[i] / 1e3, dur[i]), "warp rays-done:", [round(float(x), 1) for x in rd_t[i][w[i]]])
