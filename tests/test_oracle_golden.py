"""Pins the oracle to the reference CUDA build: tests/golden/*.npz were produced by
oracle/gen_golden_gpu.py running the unmodified reference (baseline/_ref) on a B200."""
import hashlib

import numpy as np
import pytest

from conftest import golden_bins, load_golden


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["ref_sampling_ball", "ref_sampling_frag", "ref_sampling_lvl4",
                                  "ref_sampling_lvl4_tminmax"])
def test_sampling_bit_exact(orc, name):
    z = load_golden(name)
    kw = dict(zip([str(s) for s in z["kw_names"]], [float(v) for v in z["kw_vals"]]))
    if "in_t_min" in z:
        kw["t_min"], kw["t_max"] = z["in_t_min"], z["in_t_max"]
    ri, ts, te, pi = orc.occgrid_sampling(z["rays_o"], z["rays_d"], golden_bins(z), z["aabbs"], **kw)
    assert len(ri) == int(z["n_samples"])
    np.testing.assert_array_equal(pi, z["packed_info"])
    np.testing.assert_array_equal(ts, z["t_starts"])
    np.testing.assert_array_equal(te, z["t_ends"])
    assert _sha(ri) == str(z["ray_indices_sha"])


def test_traverse_grids_bit_exact(orc):
    z = load_golden("ref_traverse_lvl4")
    iv, sm, term = orc.traverse_grids(z["rays_o"], z["rays_d"], golden_bins(z), z["aabbs"], step_size=float(z["step_size"]))
    np.testing.assert_array_equal(iv["vals"], z["iv_vals"])
    np.testing.assert_array_equal(np.packbits(iv["is_left"]), z["iv_left"])
    np.testing.assert_array_equal(np.packbits(iv["is_right"]), z["iv_right"])
    np.testing.assert_array_equal(iv["packed_info"], z["iv_packed_info"])
    assert _sha(iv["ray_indices"]) == str(z["iv_ray_sha"])
    np.testing.assert_array_equal(sm["vals"], z["sm_vals"])
    np.testing.assert_array_equal(sm["packed_info"], z["sm_packed_info"])
    assert _sha(sm["ray_indices"]) == str(z["sm_ray_sha"])
    defined = ~np.isnan(term)  # rays without samples: undefined in the reference (torch.empty)
    np.testing.assert_array_equal(term[defined], z["terminate"][defined])


def test_ray_aabb_bit_exact(orc):
    z = load_golden("ref_ray_aabb")
    tm, tM, h = orc.ray_aabb_intersect(z["rays_o"], z["rays_d"], z["aabbs"])
    np.testing.assert_array_equal(tm, z["t_mins"])
    np.testing.assert_array_equal(tM, z["t_maxs"])
    np.testing.assert_array_equal(h, z["hits"])


def test_rendering_within_tolerance(orc):
    """north_star tolerance: 1e-5 abs on weights / colours (the reference itself sums with
    atomics and a look-back scan, so it is not bit-reproducible)."""
    z = load_golden("ref_render_ball")
    o = orc.composite(z["t_starts"], z["t_ends"], z["sigmas"], z["rgbs"], packed_info=z["packed_info"],
                      render_bkgd=z["bkgd"])
    for k in ["weights", "trans", "alphas", "colors", "opacities", "depths"]:
        np.testing.assert_allclose(o[k], z[k], atol=1e-5, rtol=0, err_msg=k)
    gs, gr = orc.composite_backward(z["t_starts"], z["t_ends"], z["sigmas"], z["rgbs"], z["packed_info"], gC=z["gC"],
                                    gO=z["gO"].ravel(), gD=z["gD"].ravel(), render_bkgd=z["bkgd"])
    np.testing.assert_allclose(gs, z["g_sigmas"], atol=1e-6, rtol=1e-4)
    np.testing.assert_allclose(gr, z["g_rgbs"], atol=1e-6, rtol=1e-4)
    x = load_golden("ref_render_extras")
    w, T = orc.render_weight_from_alpha(x["alphas_in"], packed_info=z["packed_info"])
    np.testing.assert_allclose(w, x["weights_a"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(T, x["trans_a"], atol=1e-5, rtol=0)
    gs2, _ = orc.composite_backward(z["t_starts"], z["t_ends"], z["sigmas"], None, z["packed_info"], gW=x["gW"],
                                    gT=x["gT"], gA=x["gA"])
    np.testing.assert_allclose(gs2, x["g_sigmas"], atol=1e-5, rtol=1e-4)


def test_scans_within_tolerance(orc):
    s = load_golden("ref_scans")
    gy = (np.arange(len(s["x"])) % 7 + 1).astype(np.float32)
    for nm in ["inclusive_sum", "exclusive_sum", "inclusive_prod", "exclusive_prod"]:
        y = getattr(orc, nm)(s["x"], packed_info=s["packed_info"])
        for mode in ["packed", "key"]:
            np.testing.assert_allclose(y, s[f"{nm}_{mode}"], rtol=2e-5, atol=1e-6, err_msg=f"{nm}/{mode}")
        if nm.endswith("sum"):
            g = getattr(orc, nm)(gy, packed_info=s["packed_info"], reverse=True)
        else:
            g = orc.prod_backward(s["x"], y, gy, packed_info=s["packed_info"], inclusive=nm.startswith("inclusive"))
        for mode in ["packed", "key"]:
            np.testing.assert_allclose(g, s[f"{nm}_{mode}_grad"], rtol=5e-5, atol=1e-5, err_msg=f"{nm}/{mode}/grad")
