"""Golden vectors for the PDF path from the UNMODIFIED reference (baseline/_ref) on a GPU.

Run on the GPU box:  python oracle/gen_golden_pdf_gpu.py   ->  gpurun_out/golden/ref_pdf.npz, ref_propnet.npz
(copied to tests/golden/ afterwards).  Covers importance_sampling (batched + flattened input, plain and
stratified with the torch generator state recorded), searchsorted (batched, flattened) and an end-to-end
PropNetEstimator.sampling with analytic proposal densities.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import nerfacc as ref  # noqa: E402  (the reference)
from nerfacc.data_specs import RayIntervals  # noqa: E402
from nerfacc.estimators.prop_net import PropNetEstimator  # noqa: E402
from nerfacc.pdf import importance_sampling, searchsorted  # noqa: E402

assert "baseline/_ref" in ref.__file__, ref.__file__
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
torch.cuda.init()
SEED = 20240917
rng = np.random.default_rng(11)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def gen_offset():
    return int(torch.cuda.default_generators[0].get_offset())


out = {"seed": SEED, "n_out": 24}
torch.manual_seed(SEED)

# batched input: 37 rays x 65 edges, with a flat CDF row and a plateau row
b_vals = np.sort(rng.uniform(0, 4, (37, 65)).astype(np.float32), -1)
b_cdfs = np.sort(rng.uniform(0, 1, (37, 65)).astype(np.float32), -1)
b_cdfs[3] = b_cdfs[3, 0]
b_cdfs[4, 1:-1] = b_cdfs[4, 1]
out.update(b_vals=b_vals, b_cdfs=b_cdfs)
for strat in (0, 1):
    out[f"b_offset{strat}"] = gen_offset()
    iv, sm = importance_sampling(RayIntervals(vals=T(b_vals)), T(b_cdfs), 24, bool(strat))
    out[f"b_edges{strat}"], out[f"b_samples{strat}"] = N(iv.vals), N(sm.vals)

# flattened input: 29 rays with 2..40 edges
cnt = rng.integers(2, 41, 29)
f_packed = np.stack([np.cumsum(cnt) - cnt, cnt], -1).astype(np.int64)
f_vals = np.concatenate([np.sort(rng.uniform(0, 6, c)) for c in cnt]).astype(np.float32)
f_cdfs = np.concatenate([np.sort(rng.uniform(0, 1, c)) for c in cnt]).astype(np.float32)
out.update(f_vals=f_vals, f_cdfs=f_cdfs, f_packed=f_packed)
for strat in (0, 1):
    out[f"f_offset{strat}"] = gen_offset()
    iv, sm = importance_sampling(RayIntervals(vals=T(f_vals), packed_info=T(f_packed)), T(f_cdfs), 24, bool(strat))
    out[f"f_edges{strat}"], out[f"f_samples{strat}"] = N(iv.vals), N(sm.vals)

# searchsorted: batched key/query, and flattened key/query (query without ray ids -> chunk search)
ss_key = np.sort(rng.uniform(0, 1, (10, 101)).astype(np.float32), -1)
ss_query = rng.uniform(-0.1, 1.1, (10, 77)).astype(np.float32)
left, right = searchsorted(RayIntervals(vals=T(ss_key)), RayIntervals(vals=T(ss_query)))
out.update(ss_key=ss_key, ss_query=ss_query, ss_left=N(left), ss_right=N(right))
qc = rng.integers(0, 9, 29)
ssf_qpacked = np.stack([np.cumsum(qc) - qc, qc], -1).astype(np.int64)
ssf_query = rng.uniform(-0.5, 6.5, int(qc.sum())).astype(np.float32)
left, right = searchsorted(RayIntervals(vals=T(f_vals), packed_info=T(f_packed)),
                           RayIntervals(vals=T(ssf_query), packed_info=T(ssf_qpacked)))
out.update(ssf_query=ssf_query, ssf_qpacked=ssf_qpacked, ssf_left=N(left), ssf_right=N(right))
np.savez_compressed(os.path.join(OUT, "ref_pdf.npz"), **out)
print("ref_pdf.npz", {k: np.shape(v) for k, v in out.items()})


# ---- PropNetEstimator.sampling end to end, analytic proposal densities (see tests/test_gpu_pdf.py::_prop_fns)
def prop_fn(center, width, amp):
    def fn(t_starts, t_ends):
        mid = (t_starts + t_ends) * 0.5
        return amp * torch.exp(-((mid - center) / width) ** 2)
    return fn


pn = {}
for kind, near, far in [("lindisp", 0.2, 50.0), ("uniform", 2.0, 6.0)]:
    for strat in (0, 1):
        torch.manual_seed(SEED + 1)
        est = PropNetEstimator().to(dev)
        off0 = gen_offset()
        ts, te = est.sampling([prop_fn(3.0, 1.0, 4.0), prop_fn(3.2, 0.5, 8.0)], [64, 32], 16, n_rays=53,
                              near_plane=near, far_plane=far, sampling_type=kind, stratified=bool(strat))
        pn[f"{kind}_{strat}_t_starts"], pn[f"{kind}_{strat}_t_ends"] = N(ts), N(te)
        pn[f"{kind}_{strat}_offset0"], pn[f"{kind}_{strat}_offset1"] = off0, gen_offset()
pn["seed"] = SEED + 1
np.savez_compressed(os.path.join(OUT, "ref_propnet.npz"), **pn)
print("ref_propnet.npz", {k: np.shape(v) for k, v in pn.items()})
