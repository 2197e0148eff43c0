"""Config-4 importance sampling launches for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfacc_b200.data_specs import RayIntervals
from nerfacc_b200.pdf import importance_sampling
dev = torch.device("cuda:0")
R = 262144
torch.manual_seed(3)
edges = torch.sort(torch.rand(R, 65, device=dev), -1)[0]
edges[:, 0], edges[:, -1] = 0.0, 1.0
w = torch.rand(R, 64, device=dev) ** 4 + 1e-3
cdfs = torch.cat([torch.zeros(R, 1, device=dev), torch.cumsum(w, -1)], -1)
cdfs = (cdfs / cdfs[:, -1:]).contiguous()
for strat in (False, True, False, True):
    importance_sampling(RayIntervals(vals=edges), cdfs, 32, strat)
torch.cuda.synchronize()
