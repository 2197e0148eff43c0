"""Host-side mirror of the reference API: signatures, buffers, batched (CPU) maths, error behaviour."""
import inspect

import numpy as np
import pytest
import torch

import nerfacc_b200 as nfa


def _params(fn):
    return [(p.name, p.default if p.default is not inspect._empty else "<req>") for p in inspect.signature(fn).parameters.values()]


def test_public_signatures_match_reference():
    # /root/reference/nerfacc/estimators/occ_grid.py:86-104
    assert _params(nfa.OccGridEstimator.sampling)[1:] == [
        ("rays_o", "<req>"), ("rays_d", "<req>"), ("sigma_fn", None), ("alpha_fn", None), ("near_plane", 0.0),
        ("far_plane", 1e10), ("t_min", None), ("t_max", None), ("render_step_size", 1e-3), ("early_stop_eps", 1e-4),
        ("alpha_thre", 0.0), ("stratified", False), ("cone_angle", 0.0)]
    # volrend.py:15-27
    assert _params(nfa.rendering) == [
        ("t_starts", "<req>"), ("t_ends", "<req>"), ("ray_indices", None), ("n_rays", None), ("rgb_sigma_fn", None),
        ("rgb_alpha_fn", None), ("render_bkgd", None), ("expected_depths", True)]
    # volrend.py:326-334, 167-173, 497-502; pack.py:11; scan.py:14-18; grid.py:94-113
    assert [n for n, _ in _params(nfa.render_weight_from_density)] == [
        "t_starts", "t_ends", "sigmas", "packed_info", "ray_indices", "n_rays", "prefix_trans"]
    assert [n for n, _ in _params(nfa.render_transmittance_from_alpha)] == [
        "alphas", "packed_info", "ray_indices", "n_rays", "prefix_trans"]
    assert [n for n, _ in _params(nfa.accumulate_along_rays)] == ["weights", "values", "ray_indices", "n_rays"]
    assert _params(nfa.pack_info) == [("ray_indices", "<req>"), ("n_rays", None)]
    for f in (nfa.inclusive_sum, nfa.exclusive_sum, nfa.inclusive_prod, nfa.exclusive_prod):
        assert _params(f) == [("inputs", "<req>"), ("packed_info", None), ("indices", None)]
    assert [n for n, _ in _params(nfa.traverse_grids)] == [
        "rays_o", "rays_d", "binaries", "aabbs", "near_planes", "far_planes", "step_size", "cone_angle",
        "traverse_steps_limit", "over_allocate", "rays_mask", "t_sorted", "t_indices", "hits"]
    # pdf.py:12-15, :64-69; estimators/prop_net.py:37-52, :131, :156-161, :196-198; losses.py:7-13
    assert _params(nfa.searchsorted) == [("sorted_sequence", "<req>"), ("values", "<req>")]
    assert _params(nfa.importance_sampling) == [("intervals", "<req>"), ("cdfs", "<req>"),
                                                ("n_intervals_per_ray", "<req>"), ("stratified", False)]
    assert _params(nfa.PropNetEstimator.sampling)[1:] == [
        ("prop_sigma_fns", "<req>"), ("prop_samples", "<req>"), ("num_samples", "<req>"), ("n_rays", "<req>"),
        ("near_plane", "<req>"), ("far_plane", "<req>"), ("sampling_type", "lindisp"), ("stratified", False),
        ("requires_grad", False)]
    assert _params(nfa.PropNetEstimator.compute_loss)[1:] == [("trans", "<req>"), ("loss_scaler", 1.0)]
    assert _params(nfa.PropNetEstimator.update_every_n_steps)[1:] == [("trans", "<req>"), ("requires_grad", False),
                                                                      ("loss_scaler", 1.0)]
    from nerfacc.estimators.prop_net import get_proposal_requires_grad_fn
    assert _params(get_proposal_requires_grad_fn) == [("target", 5.0), ("num_steps", 1000)]
    assert [n for n, _ in _params(nfa.distortion)] == ["weights", "t_starts", "t_ends", "ray_indices", "n_rays"]
    import nerfacc
    assert nerfacc.OccGridEstimator is nfa.OccGridEstimator and nerfacc.rendering is nfa.rendering
    assert nerfacc.PropNetEstimator is nfa.PropNetEstimator and nerfacc.importance_sampling is nfa.importance_sampling


def test_estimator_buffers_and_state_dict():
    est = nfa.OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=16, levels=3)
    sd = est.state_dict()
    assert sorted(sd) == ["aabbs", "binaries", "occs", "resolution"]  # reference occ_grid.py:67-83 (persistent only)
    assert sd["binaries"].shape == (3, 16, 16, 16) and sd["binaries"].dtype == torch.bool
    assert sd["occs"].shape == (3 * 16 ** 3,) and sd["resolution"].dtype == torch.int32
    np.testing.assert_allclose(sd["aabbs"].numpy(), [[-1] * 3 + [1] * 3, [-2] * 3 + [2] * 3, [-4] * 3 + [4] * 3])
    est2 = nfa.OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=16, levels=3)
    est2.load_state_dict(sd)
    with pytest.raises(ValueError):
        nfa.OccGridEstimator([-1, -1, -1, 1, 1, 1], contraction_type=1)


def test_grid_update_and_mark_invisible_cells():
    torch.manual_seed(0)
    est = nfa.OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=32, levels=4)
    K = torch.tensor([[[100.0, 0, 50.0], [0, 100.0, 50.0], [0, 0, 1]]])
    pose = torch.tensor([[[-1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5]]])
    est.mark_invisible_cells(K, pose, 100, 100)
    # exact counts of the reference test (tests/test_grid.py:207-233)
    assert int((est.occs == -1).sum()) == 77660 and int((est.occs == 0).sum()) == 53412
    est.train()
    est.update_every_n_steps(0, lambda x: (x.norm(dim=-1, keepdim=True) < 0.5).float(), occ_thre=0.01)
    assert est.binaries.dtype == torch.bool and est.binaries.any() and not est.binaries.all()
    est.eval()
    with pytest.raises(RuntimeError):
        est.update_every_n_steps(0, lambda x: x[:, :1])


def test_batched_rendering_on_cpu_matches_oracle(orc):
    """Config 1 of BASELINE.json: the batched [n_rays, n_samples] path is pure torch and runs on CPU."""
    g = torch.Generator().manual_seed(3)
    R, S = 64, 40
    ts = torch.rand(R, S, generator=g).sort(-1).values
    te = ts + torch.rand(R, S, generator=g) * 0.02
    sig = (5 * torch.rand(R, S, generator=g)).requires_grad_(True)
    rgb = torch.rand(R, S, 3, generator=g).requires_grad_(True)
    bk = torch.tensor([0.1, 0.2, 0.3])
    col, op, dep, ex = nfa.rendering(ts, te, rgb_sigma_fn=lambda a, b, c: (rgb, sig), render_bkgd=bk)
    pi = np.stack([np.arange(R) * S, np.full(R, S)], -1)
    o = orc.composite(ts.numpy().ravel(), te.numpy().ravel(), sig.detach().numpy().ravel(),
                      rgb.detach().numpy().reshape(-1, 3), packed_info=pi, render_bkgd=bk.numpy())
    np.testing.assert_allclose(col.detach().numpy(), o["colors"], atol=1e-5)
    np.testing.assert_allclose(op.detach().numpy(), o["opacities"], atol=1e-5)
    np.testing.assert_allclose(dep.detach().numpy(), o["depths"], atol=1e-5)
    np.testing.assert_allclose(ex["weights"].detach().numpy().ravel(), o["weights"], atol=1e-5)
    gC = torch.rand(R, 3, generator=g)
    (col * gC).sum().backward()
    gs, gr = orc.composite_backward(ts.numpy().ravel(), te.numpy().ravel(), sig.detach().numpy().ravel(),
                                    rgb.detach().numpy().reshape(-1, 3), pi, gC=gC.numpy(), render_bkgd=bk.numpy())
    np.testing.assert_allclose(sig.grad.numpy().ravel(), gs, atol=1e-5)
    np.testing.assert_allclose(rgb.grad.numpy().reshape(-1, 3), gr, atol=1e-5)
    # alpha route, batched
    al = torch.rand(R, S, generator=g) * 0.3
    w, T = nfa.render_weight_from_alpha(al)
    ow, oT = orc.render_weight_from_alpha(al.numpy().ravel(), packed_info=pi)
    np.testing.assert_allclose(w.numpy().ravel(), ow, atol=1e-6)
    np.testing.assert_allclose(T.numpy().ravel(), oT, atol=1e-6)
    vis = nfa.render_visibility_from_alpha(al, early_stop_eps=0.5, alpha_thre=0.1)
    np.testing.assert_array_equal(vis.numpy().ravel(), (oT >= 0.5) & (al.numpy().ravel() >= 0.1))


def test_batched_scans_and_errors_on_cpu():
    x = torch.arange(1.0, 7.0).reshape(2, 3)
    assert nfa.inclusive_sum(x).tolist() == [[1, 3, 6], [4, 9, 15]]
    assert nfa.exclusive_sum(x).tolist() == [[0, 1, 3], [0, 4, 9]]
    assert nfa.inclusive_prod(x).tolist() == [[1, 2, 6], [4, 20, 120]]
    assert nfa.exclusive_prod(x).tolist() == [[1, 1, 2], [1, 4, 20]]
    flat = x.flatten()
    pi = torch.tensor([[0, 3], [3, 3]])
    with pytest.raises(ValueError):  # reference scan.py:47-50
        nfa.inclusive_sum(flat, packed_info=pi, indices=torch.zeros(6, dtype=torch.long))
    with pytest.raises(NotImplementedError):  # packed path is CUDA-only, no CPU fallback (reference pack.py:47-48)
        nfa.pack_info(torch.tensor([0, 0, 1]))
    with pytest.raises(NotImplementedError):
        nfa.inclusive_sum(flat, packed_info=pi)
    with pytest.raises(ValueError):  # reference volrend.py:84-87
        nfa.rendering(flat, flat)
    acc = nfa.accumulate_along_rays(torch.ones(2, 3), torch.ones(2, 3, 2))
    assert acc.shape == (2, 2) and acc[0, 0] == 3
