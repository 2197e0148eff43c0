"""CPU checks of the PDF path (importance_sampling / searchsorted / PropNetEstimator host logic).

* oracle vs the reference's docstring known answers (pdf.py:41-57, :109-122) and vs the torch
  cross-check implementation the reference's own test uses (tests/test_pdf.py:62-91);
* product headers (csrc/pdf.cuh through tests/host_sim) bit-exact vs the oracle;
* the pure-torch pieces of nerfacc_b200.pdf / estimators.prop_net.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

F = C.POINTER(C.c_float)
I64 = C.POINTER(C.c_int64)
U8 = C.POINTER(C.c_uint8)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _sorted_rows(rng, n_rays, n_edges, lo=0.0, hi=1.0):
    return np.sort(rng.uniform(lo, hi, (n_rays, n_edges)).astype(np.float32), -1)


def _sim_is(lib, vals, cdfs, n, packed=None, cnts=None, stratified=False, seed=0, offset=0, stot=None):
    vals, cdfs = np.ascontiguousarray(vals, np.float32), np.ascontiguousarray(cdfs, np.float32)
    if packed is None:
        n_rays, in_edges = vals.shape[0], vals.shape[1]
    else:
        packed = np.ascontiguousarray(packed, np.int64)
        n_rays, in_edges = packed.shape[0], 0
    if cnts is None:
        s = np.zeros((n_rays, n), np.float32)
        e = np.zeros((n_rays, n + 1), np.float32)
        ts = te = None
        smin = smax = 0.0
        lind = 0
        if stot is not None:
            smin, smax, lind = stot
            ts, te = np.zeros_like(s), np.zeros_like(s)
        lib.sim_importance_sampling(C.c_int32(n_rays), _p(vals, F), _p(cdfs, F), _p(packed, I64), C.c_int64(in_edges),
                                    None, None, C.c_int64(n), C.c_int32(stratified), C.c_uint64(seed),
                                    C.c_uint64(offset), _p(s, F), None, _p(e, F), None, None, None, _p(ts, F),
                                    _p(te, F), C.c_float(smin), C.c_float(smax), C.c_int32(lind))
        return e, s, ts, te
    cnts = np.ascontiguousarray(cnts, np.int64)
    sp = np.stack([np.cumsum(cnts) - cnts, cnts], -1).astype(np.int64)
    ec = (cnts + 1) * (cnts > 0)
    ep = np.stack([np.cumsum(ec) - ec, ec], -1).astype(np.int64)
    ns, ne = int(cnts.sum()), int(ec.sum())
    s, sr = np.zeros(ns, np.float32), np.zeros(ns, np.int64)
    e, er = np.zeros(ne, np.float32), np.zeros(ne, np.int64)
    el, erg = np.zeros(ne, np.uint8), np.zeros(ne, np.uint8)
    lib.sim_importance_sampling(C.c_int32(n_rays), _p(vals, F), _p(cdfs, F), _p(packed, I64), C.c_int64(in_edges),
                                _p(sp, I64), _p(ep, I64), C.c_int64(0), C.c_int32(stratified), C.c_uint64(seed),
                                C.c_uint64(offset), _p(s, F), _p(sr, I64), _p(e, F), _p(er, I64), _p(el, U8), _p(erg, U8),
                                None, None, C.c_float(0), C.c_float(0), C.c_int32(0))
    return dict(vals=e, packed_info=ep, ray_indices=er, is_left=el.astype(bool), is_right=erg.astype(bool)), \
        dict(vals=s, packed_info=sp, ray_indices=sr)


# --------------------------------------------------------------------------------------
# oracle: known answers
# --------------------------------------------------------------------------------------

def test_oracle_docstring_kats(orc):
    # pdf.py:109-122
    iv, sm = orc.importance_sampling(np.array([0.0, 1.0, 0.0, 1.0, 2.0]), np.array([0.0, 0.5, 0.0, 0.5, 1.0]), 2,
                                     packed_info=np.array([[0, 2], [2, 3]]))
    assert np.array_equal(iv, np.array([[0.0, 0.5, 1.0], [0.0, 1.0, 2.0]], np.float32))
    assert np.array_equal(sm, np.array([[0.25, 0.75], [0.5, 1.5]], np.float32))
    # pdf.py:41-57
    left, right = orc.searchsorted(np.array([0.0, 1.0, 0.0, 1.0, 2.0]), np.array([0.5, 1.5, 2.5]),
                                   key_packed_info=np.array([[0, 2], [2, 3]]),
                                   query_packed_info=np.array([[0, 1], [1, 2]]))
    assert left.tolist() == [0, 3, 3] and right.tolist() == [1, 4, 4]


def test_oracle_philox_known_values(orc):
    # torch.manual_seed(42); torch.rand(3, device="cuda") -> 0.6130, 0.0101, 0.3984 (Philox4x32-10, one
    # subsequence per element, first output word): the same stream position the reference's jitter uses.
    got = [orc.philox_uniform(42, i, 0) for i in range(3)]
    assert np.allclose(got, [0.6130, 0.0101, 0.3984], atol=5e-5)
    # Random123 known-answer test vector for philox4x32-10: counter = key = 0
    fn = orc.lib().orc_philox_word
    fn.restype = C.c_uint32
    assert fn(C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)) == 0x6627E8D5


def test_oracle_matches_reference_torch_crosscheck(orc):
    """The reference's own test (tests/test_pdf.py:62-91): CUDA result vs `_sample_from_weighted`, atol 1e-4."""
    from nerfacc_b200.pdf import _sample_from_weighted
    rng = np.random.default_rng(42)
    vals, cdfs = _sorted_rows(rng, 5, 101), _sorted_rows(rng, 5, 101)
    iv, sm = orc.importance_sampling(vals, cdfs, 100)
    for i in range(5):
        b, c = torch.from_numpy(vals[i:i + 1]), torch.from_numpy(cdfs[i:i + 1])
        e_ref, m_ref = _sample_from_weighted(b, c[:, 1:] - c[:, :-1], 100, False, b.min(), b.max())
        assert np.allclose(iv[i:i + 1], e_ref.numpy(), atol=1e-4)
        assert np.allclose(sm[i:i + 1], m_ref.numpy(), atol=1e-4)


def test_oracle_searchsorted_matches_torch(orc):
    """tests/test_pdf.py:44-60."""
    rng = np.random.default_rng(1)
    key, query = _sorted_rows(rng, 10, 101), _sorted_rows(rng, 10, 101)
    left, right = orc.searchsorted(key, query)
    want = torch.searchsorted(torch.from_numpy(key), torch.from_numpy(query), right=True).clamp(0, 100).numpy()
    assert np.array_equal(right, want)
    kv = np.take_along_axis(key, left, -1)
    kr = np.take_along_axis(key, right, -1)
    inside = (query >= key[:, :1]) & (query < key[:, -1:])
    assert np.all((kv <= query)[inside]) and np.all((query < kr)[inside])


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "ref_pdf.npz")), reason="ref_pdf golden not generated yet")
def test_oracle_matches_reference_cuda_goldens(orc):
    z = np.load(os.path.join(GOLDEN, "ref_pdf.npz"))
    for tag in ("b", "f"):
        packed = z[f"{tag}_packed"] if f"{tag}_packed" in z else None
        for strat in (0, 1):
            iv, sm = orc.importance_sampling(z[f"{tag}_vals"], z[f"{tag}_cdfs"], int(z["n_out"]), packed_info=packed,
                                             stratified=bool(strat), seed=int(z["seed"]),
                                             offset=int(z[f"{tag}_offset{strat}"]))
            assert np.array_equal(sm.view(np.uint32), z[f"{tag}_samples{strat}"].view(np.uint32))
            assert np.array_equal(iv.view(np.uint32), z[f"{tag}_edges{strat}"].view(np.uint32))
    left, right = orc.searchsorted(z["ss_key"], z["ss_query"])
    assert np.array_equal(left, z["ss_left"]) and np.array_equal(right, z["ss_right"])
    left, right = orc.searchsorted(z["f_vals"], z["ssf_query"], key_packed_info=z["f_packed"],
                                   query_packed_info=z["ssf_qpacked"])
    assert np.array_equal(left, z["ssf_left"]) and np.array_equal(right, z["ssf_right"])


# --------------------------------------------------------------------------------------
# product headers on the CPU vs the oracle (bit-exact)
# --------------------------------------------------------------------------------------

@pytest.mark.parametrize("stratified", [False, True])
def test_header_importance_sampling_batched(host_sim, orc, stratified):
    rng = np.random.default_rng(3)
    for n_edges, n_out in [(2, 1), (2, 7), (33, 64), (257, 96), (101, 100), (5, 2)]:
        vals, cdfs = _sorted_rows(rng, 17, n_edges, 0.0, 3.0), _sorted_rows(rng, 17, n_edges)
        cdfs[3] = cdfs[3, 0]          # flat CDF: every interval hits the `< 1e-10` branch
        cdfs[4, 1:-1] = cdfs[4, 1]    # plateaus (repeated CDF values)
        e0, s0 = orc.importance_sampling(vals, cdfs, n_out, stratified=stratified, seed=1234, offset=8)
        e1, s1, _, _ = _sim_is(host_sim, vals, cdfs, n_out, stratified=stratified, seed=1234, offset=8)
        assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32))
        assert np.array_equal(e0.view(np.uint32), e1.view(np.uint32))
        assert np.all(np.diff(e1, axis=-1) >= 0)


def test_header_importance_sampling_flattened(host_sim, orc):
    rng = np.random.default_rng(4)
    cnt_in = rng.integers(2, 40, 23)
    packed = np.stack([np.cumsum(cnt_in) - cnt_in, cnt_in], -1)
    vals = np.concatenate([np.sort(rng.uniform(0, 5, c)) for c in cnt_in]).astype(np.float32)
    cdfs = np.concatenate([np.sort(rng.uniform(0, 1, c)) for c in cnt_in]).astype(np.float32)
    # flattened in, batched out
    e0, s0 = orc.importance_sampling(vals, cdfs, 16, packed_info=packed)
    e1, s1, _, _ = _sim_is(host_sim, vals, cdfs, 16, packed=packed)
    assert np.array_equal(e0, e1) and np.array_equal(s0, s1)
    # flattened in, per-ray counts out (some rays get none)
    cnts = rng.integers(0, 12, 23)
    cnts[5] = 0
    cnts[7] = 1
    iv0, sm0 = orc.importance_sampling(vals, cdfs, cnts, packed_info=packed, stratified=True, seed=7, offset=4)
    iv1, sm1 = _sim_is(host_sim, vals, cdfs, 0, packed=packed, cnts=cnts, stratified=True, seed=7, offset=4)
    for k in iv0:
        assert np.array_equal(iv0[k], iv1[k]), k
    for k in sm0:
        assert np.array_equal(sm0[k], sm1[k]), k
    assert iv0["vals"].size == cnts.sum() + (cnts > 0).sum()
    assert iv0["is_left"].sum() == cnts.sum() == iv0["is_right"].sum()


def test_header_philox_and_searchsorted(host_sim, orc):
    for seed, sub, off in [(0, 0, 0), (42, 0, 0), (42, 5, 4), (2 ** 40 + 17, 2 ** 33 + 5, 2 ** 36 + 8), (7, 3, 13)]:
        assert host_sim.sim_philox_uniform(seed, sub, off) == orc.philox_uniform(seed, sub, off)
    rng = np.random.default_rng(5)
    key, query = _sorted_rows(rng, 9, 50), rng.uniform(-0.2, 1.2, (9, 31)).astype(np.float32)
    l0, r0 = orc.searchsorted(key, query)
    l1, r1 = np.zeros_like(l0), np.zeros_like(r0)
    host_sim.sim_searchsorted(C.c_int64(query.size), _p(query, F), None, None, C.c_int32(9), C.c_int64(31), _p(key, F),
                              None, C.c_int64(50), _p(l1, I64), _p(r1, I64))
    assert np.array_equal(l0, l1) and np.array_equal(r0, r1)
    # flattened query without ray ids (chunk search) against flattened keys
    kc, qc = rng.integers(1, 30, 12), rng.integers(0, 9, 12)
    kp = np.stack([np.cumsum(kc) - kc, kc], -1).astype(np.int64)
    qp = np.stack([np.cumsum(qc) - qc, qc], -1).astype(np.int64)
    kv = np.concatenate([np.sort(rng.uniform(0, 1, c)) for c in kc]).astype(np.float32)
    qv = rng.uniform(-0.1, 1.1, int(qc.sum())).astype(np.float32)
    l0, r0 = orc.searchsorted(kv, qv, key_packed_info=kp, query_packed_info=qp)
    l1, r1 = np.zeros_like(l0), np.zeros_like(r0)
    host_sim.sim_searchsorted(C.c_int64(qv.size), _p(qv, F), _p(qp, I64), None, C.c_int32(12), C.c_int64(0), _p(kv, F),
                              _p(kp, I64), C.c_int64(0), _p(l1, I64), _p(r1, I64))
    assert np.array_equal(l0, l1) and np.array_equal(r0, r1)
    ray_of_q = np.repeat(np.arange(12), qc)
    assert np.all(l0 >= kp[ray_of_q, 0]) and np.all(r0 < kp[ray_of_q].sum(-1))


def test_header_stot_matches_torch(host_sim):
    """The fused s -> t mapping reproduces prop_net._transform_stot op for op."""
    from nerfacc_b200.estimators.prop_net import _stot_constants, _transform_stot
    rng = np.random.default_rng(6)
    vals = np.stack([np.zeros(9, np.float32), np.ones(9, np.float32)], -1)
    cdfs = vals.copy()
    for kind, near, far in [("lindisp", 0.2, 1e3), ("uniform", 0.05, 6.0), ("lindisp", 2.0, 6.0)]:
        smin, smax, lind = _stot_constants(kind, near, far)
        e, s, ts, te = _sim_is(host_sim, vals, cdfs, 48, stot=(smin, smax, int(lind)))
        want = _transform_stot(kind, torch.from_numpy(e), near, far).numpy()
        assert np.array_equal(ts, want[:, :-1]) and np.array_equal(te, want[:, 1:])
    assert rng is not None


# --------------------------------------------------------------------------------------
# torch-side host logic
# --------------------------------------------------------------------------------------

def test_pdf_loss_matches_lossfun_outer(orc):
    """tests/test_pdf.py:94-127 with the native searchsorted replaced by the oracle's."""
    import nerfacc_b200.estimators.prop_net as pn
    from nerfacc_b200.data_specs import RayIntervals
    rng = np.random.default_rng(42)
    vals, cdfs = _sorted_rows(rng, 5, 101), _sorted_rows(rng, 5, 101)
    e, _ = orc.importance_sampling(vals, cdfs, 10)
    # the two formulations agree where the envelope spans the query range (outside it `searchsorted` clips both
    # ids to the same edge while `_outer` still credits the first/last envelope bin)
    e[:, 0], e[:, -1] = vals[:, 0], vals[:, -1]
    cdfs2 = _sorted_rows(rng, 5, 11)

    def cpu_searchsorted(key, query):
        left, right = orc.searchsorted(key.vals.numpy(), query.vals.numpy())
        return torch.from_numpy(left), torch.from_numpy(right)

    saved = pn.searchsorted
    pn.searchsorted = cpu_searchsorted
    try:
        loss = pn._pdf_loss(RayIntervals(torch.from_numpy(vals)), torch.from_numpy(cdfs),
                            RayIntervals(torch.from_numpy(e)), torch.from_numpy(cdfs2))
    finally:
        pn.searchsorted = saved
    t, c = torch.from_numpy(vals), torch.from_numpy(cdfs)
    te, ce = torch.from_numpy(e), torch.from_numpy(cdfs2)
    loss2 = pn._lossfun_outer(t, c[:, 1:] - c[:, :-1], te, ce[:, 1:] - ce[:, :-1])
    assert torch.allclose(loss, loss2, atol=1e-4)


def test_proposal_requires_grad_schedule():
    from nerfacc_b200.estimators.prop_net import get_proposal_requires_grad_fn
    fn = get_proposal_requires_grad_fn(target=5.0, num_steps=1000)
    fired = [fn(s) for s in range(3000)]
    assert sum(fired[:10]) >= 4            # nearly every step at the start
    late = np.flatnonzero(fired[2000:])
    assert np.all(np.diff(late) == 6)      # every target+1 steps once the schedule saturates


def test_native_entry_points_need_cuda():
    import nerfacc_b200 as nfa
    iv = nfa.RayIntervals(vals=torch.rand(4, 9).sort(-1)[0])
    with pytest.raises(NotImplementedError):
        nfa.importance_sampling(iv, torch.rand(4, 9).sort(-1)[0], 8)
    with pytest.raises(NotImplementedError):
        nfa.searchsorted(iv, iv)
