"""Oracle vs the reference's own known-answer vectors (tests/ and docstrings).

Each case cites where the numbers come from in /root/reference.
"""
import numpy as np


def test_pack_info_kat(orc):
    # tests/test_pack.py:11-18
    assert orc.pack_info([0, 2, 2, 2, 2], 3).tolist() == [[0, 1], [1, 0], [1, 4]]
    # nerfacc/pack.py:29-32
    assert orc.pack_info([0, 0, 1, 1, 1, 2, 2, 2, 2], 3).tolist() == [[0, 2], [2, 3], [5, 4]]


def test_scan_docstring_kats(orc):
    x = np.arange(1, 10, dtype=np.float32)
    pi = np.array([[0, 2], [2, 3], [5, 4]])
    # nerfacc/scan.py:41-44, 105-108, 175-178, 240-243
    np.testing.assert_array_equal(orc.inclusive_sum(x, pi), [1, 3, 3, 7, 12, 6, 13, 21, 30])
    np.testing.assert_array_equal(orc.exclusive_sum(x, pi), [0, 1, 0, 3, 7, 0, 6, 13, 21])
    np.testing.assert_array_equal(orc.inclusive_prod(x, pi), [1, 2, 3, 12, 60, 6, 42, 336, 3024])
    np.testing.assert_array_equal(orc.exclusive_prod(x, pi), [1, 1, 1, 3, 12, 1, 6, 42, 336])
    idx = [0, 0, 1, 1, 1, 2, 2, 2, 2]
    np.testing.assert_array_equal(orc.exclusive_prod(x, indices=idx), [1, 1, 1, 3, 12, 1, 6, 42, 336])
    np.testing.assert_array_equal(orc.inclusive_sum(x, indices=idx), [1, 3, 3, 7, 12, 6, 13, 21, 30])


def test_volrend_docstring_kats(orc):
    # nerfacc/volrend.py:256-263, 361-369
    ts = np.arange(7, dtype=np.float32)
    te = ts + 1
    sg = np.array([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9], np.float32)
    ri = [0, 0, 0, 1, 1, 2, 2]
    w, T, a = orc.render_weight_from_density(ts, te, sg, ray_indices=ri, n_rays=3)
    np.testing.assert_allclose(T, [1.00, 0.67, 0.30, 1.00, 0.45, 1.00, 1.00], atol=5e-3)
    np.testing.assert_allclose(a, [0.33, 0.55, 0.095, 0.55, 0.095, 0.00, 0.59], atol=5e-3)
    np.testing.assert_allclose(w, [0.33, 0.37, 0.03, 0.55, 0.04, 0.00, 0.59], atol=5e-3)
    # nerfacc/volrend.py:198-201, 312-316
    w, T = orc.render_weight_from_alpha(sg, ray_indices=ri, n_rays=3)
    np.testing.assert_allclose(T, [1.0, 0.6, 0.12, 1.0, 0.2, 1.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(w, [0.4, 0.48, 0.012, 0.8, 0.02, 0.0, 0.9], atol=1e-6)


def test_rendering_goldens(orc):
    # tests/test_rendering.py:38-57 (weights from alpha, ray 1 empty)
    al = np.array([0.4, 0.3, 0.8, 0.8, 0.5], np.float32)
    w, _ = orc.render_weight_from_alpha(al, ray_indices=[0, 2, 2, 2, 2], n_rays=3)
    np.testing.assert_allclose(w, [0.4, 0.3, 0.7 * 0.8, 0.14 * 0.8, 0.028 * 0.5], rtol=1e-6)
    # tests/test_rendering.py:110-133 (weights_ref / sigmas_grad_ref)
    sg = np.array([0.4, 0.8, 0.1, 0.8, 0.1], np.float32)
    ts = np.random.default_rng(0).random(5).astype(np.float32)
    te = ts + 1.0
    pi = np.array([[0, 1], [1, 0], [1, 4]])
    w, _, _ = orc.render_weight_from_density(ts, te, sg, packed_info=pi)
    np.testing.assert_allclose(w, [0.3297, 0.5507, 0.0428, 0.2239, 0.0174], atol=1e-4)
    gs, _ = orc.composite_backward(ts, te, sg, None, pi, gW=np.ones(5, np.float32))
    np.testing.assert_allclose(gs, [0.6703, 0.1653, 0.1653, 0.1653, 0.1653], atol=1e-4)


def test_accumulate_with_empty_ray(orc):
    # tests/test_rendering.py:87-106
    ri = np.array([0, 2, 2, 2, 2])
    w = np.array([0.4, 0.3, 0.8, 0.8, 0.5], np.float32)
    v = np.random.default_rng(1).random((5, 2)).astype(np.float32)
    out = orc.accumulate_along_rays(w, v, ri, 3)
    np.testing.assert_allclose(out[0], w[0] * v[0], rtol=1e-6)
    assert (out[1] == 0).all()
    np.testing.assert_allclose(out[2], (w[1:, None] * v[1:]).sum(0), rtol=1e-6)


def test_visibility_goldens(orc):
    # tests/test_rendering.py:8-34
    ri = [0, 2, 2, 2, 2]
    al = np.array([0.4, 0.3, 0.8, 0.8, 0.5], np.float32)
    _, T = orc.render_weight_from_alpha(al, ray_indices=ri, n_rays=3)
    np.testing.assert_array_equal(T >= 0.03, [True, True, True, True, False])
    np.testing.assert_array_equal((T >= 0.05) & (al >= 0.35), [True, False, True, True, False])


def test_traverse_properties(orc):
    """tests/test_grid.py:39-68 and :135-159 restated for the oracle."""
    rng = np.random.default_rng(42)
    ro = rng.standard_normal((10, 3)).astype(np.float32)
    rd = rng.standard_normal((10, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    from nerfacc_b200 import scenes
    aabbs = scenes.nested_aabbs(4)
    bins = rng.random((4, 32, 32, 32)) > 0.5
    iv, sm, _ = orc.traverse_grids(ro, rd, bins, aabbs)
    ts, te = iv["vals"][iv["is_left"]], iv["vals"][iv["is_right"]]
    assert len(ts) == len(te) == len(sm["vals"]) > 0
    pos = ro[sm["ray_indices"]] + rd[sm["ray_indices"]] * ((ts + te) / 2.0)[:, None]
    # every midpoint lies in an occupied cell of its mip level
    u = (pos + 1.0) / 2.0 - 0.5
    reach = np.maximum(np.abs(u).max(-1), 0.1)
    lvl = np.maximum(np.frexp(reach)[1] + 1, 0)
    assert (lvl < 4).all()
    cell = np.minimum(((u / (2.0 ** lvl)[:, None] + 0.5) * 32).astype(np.int64), 31)
    assert bins[lvl, cell[:, 0], cell[:, 1], cell[:, 2]].all()
    # near / far planes respected within half a step
    iv, _, _ = orc.traverse_grids(np.array([[-1.0, 0, 0]], np.float32),
                                  (np.array([[1.0, 0.01, 0.01]]) / np.linalg.norm([1.0, 0.01, 0.01])).astype(np.float32),
                                  np.ones((1, 1, 1, 1), bool), np.array([[0, 0, 0, 1, 1, 1]], np.float32),
                                  near_planes=np.array([1.2], np.float32), far_planes=np.array([1.5], np.float32),
                                  step_size=0.05)
    assert (iv["vals"] >= 1.2 - 0.025).all() and (iv["vals"] <= 1.5 + 0.025).all() and len(iv["vals"]) > 0
