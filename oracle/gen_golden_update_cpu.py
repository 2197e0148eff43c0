"""Golden vectors for grid maintenance from the UNMODIFIED reference, on the CPU.

    python oracle/gen_golden_update_cpu.py        # writes tests/golden/ref_occ_update.npz

`OccGridEstimator._update` (/root/reference/nerfacc/estimators/occ_grid.py:367-404) is plain torch, so the reference
runs here without a GPU.  The closure records the points it was asked about and what it answered; together with
`occs` before and after and the resulting `binaries` that pins the oracle's restatement (oracle.c orc_occ_*) and,
through it, the CUDA kernels of nerfacc_b200/csrc/occ_update.cu.  Nothing of nerfacc_b200 is imported.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import nerfacc  # noqa: E402

assert nerfacc.__file__.startswith("/root/reference")
import nerfacc.volrend as _rv  # noqa: E402

out = {}
for case, (step, occ_thre, scale) in {"warm": (10, 0.01, 0.05), "sampled": (1000, 0.01, 0.002),
                                      "sampled_hi": (1000, 0.5, 0.9)}.items():
    torch.manual_seed(5)
    est = nerfacc.OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=[16, 12, 20], levels=2)
    est.train()
    est.occs = torch.rand_like(est.occs) * 0.02
    est.occs[::7] = -1.0                       # cells marked invisible
    est.binaries = (est.occs > 0.01).view(est.binaries.shape)
    rec = []

    def occ_eval_fn(x):
        d = torch.exp(-4.0 * (x * x).sum(-1, keepdim=True)) * scale * (1.0 + 0.5 * torch.sin(37.0 * x[:, :1]))
        rec.append((x.clone(), d.squeeze(-1).clone()))
        return d

    before = est.occs.clone()
    est._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=0.95, warmup_steps=256)
    res = est.resolution.float()
    ids, occ = [], []
    for lvl, (x, d) in enumerate(rec):
        lo, hi = est.aabbs[lvl, :3], est.aabbs[lvl, 3:]
        c = torch.floor((x - lo) / (hi - lo) * res).long()
        c = torch.minimum(torch.clamp(c, min=0), est.resolution.long() - 1)
        ids.append(lvl * est.cells_per_lvl + (c[:, 0] * est.resolution[1] + c[:, 1]) * est.resolution[2] + c[:, 2])
        occ.append(d)
    out[case + "_occs_before"] = before.numpy()
    out[case + "_ids"] = torch.cat(ids).numpy()
    out[case + "_occ"] = torch.cat(occ).numpy()
    out[case + "_occs_after"] = est.occs.numpy().copy()
    out[case + "_binaries"] = est.binaries.numpy().copy()
    out[case + "_args"] = np.array([occ_thre, 0.95], np.float32)
    print(case, "cells updated", len(out[case + "_ids"]), "unique", len(np.unique(out[case + "_ids"])), "occupied",
          int(est.binaries.sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_occ_update.npz"), **out)
