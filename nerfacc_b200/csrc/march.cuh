// march.cuh -- per-ray occupancy-grid traversal, constant-step fast path.
//
// Restates the semantics of the reference's traverse_grids_kernel
// (/root/reference/nerfacc/cuda/csrc/grid.cu:68-282 with the helpers in
// include/utils_grid.cuh:10-142) for cone_angle == 0 and step_size > 0, but
// re-organised for a one-pass, no-per-sample-work march:
//
//  * the DDA walks cells exactly like the reference (same f32 evaluation order,
//    same FMA contraction as the reference's sm_100a SASS -- see DESIGN.md);
//  * the marching variable is never stepped sample by sample.  Because all
//    sample starts lie on one lattice (lattice.cuh), an occupied stretch of
//    cells is turned into a *run* (t_first, n) with two closed-form seeks, and
//    empty cells only raise a pending skip target;
//  * the per-sample arrays are produced later by the expand kernel from the runs.
//
// The code is host+device so that tests/host_sim can run the very same logic on
// the CPU against the oracle (oracle/oracle.c) without a GPU.
#pragma once

#include "lattice.cuh"

namespace nfa {

// ---------------------------------------------------------------------------
// Bit-packed occupancy: 4x4x4-cell bricks, one uint64 per brick, plus a
// 1-bit-per-brick "any cell occupied" mip.  Derived cache of the estimator's
// bool `binaries` (reference estimators/occ_grid.py:73-76).
// ---------------------------------------------------------------------------
struct OccGeom {
    int res[3];   // cells per axis
    int nb[3];    // bricks per axis = ceil(res/4)
    int wpl;      // words (bricks) per level
    int n_grids;
};

NFA_HD OccGeom occ_geom(int n_grids, int rx, int ry, int rz)
{
    OccGeom g;
    g.res[0] = rx; g.res[1] = ry; g.res[2] = rz;
    g.nb[0] = (rx + 3) >> 2; g.nb[1] = (ry + 3) >> 2; g.nb[2] = (rz + 3) >> 2;
    g.wpl = g.nb[0] * g.nb[1] * g.nb[2];
    g.n_grids = n_grids;
    return g;
}

struct OccView {
    const uint64_t* words;   // [n_grids * wpl]
    const uint32_t* coarse;  // [(n_grids * wpl + 31) / 32]; may live in shared memory
    OccGeom g;
};

// Per-thread cursor with a one-brick register cache.
struct OccCursor {
    int brick;      // cached brick word index, -1 = none
    uint64_t word;
    NFA_HD void reset() { brick = -1; word = 0; }
    NFA_HD bool test(const OccView& v, int level, int ix, int iy, int iz)
    {
        const int b = ((ix >> 2) * v.g.nb[1] + (iy >> 2)) * v.g.nb[2] + (iz >> 2) + level * v.g.wpl;
        if (b != brick) {
            brick = b;
            const uint32_t c = v.coarse[b >> 5];
            word = ((c >> (b & 31)) & 1u) ? v.words[b] : 0ull;
        }
        const int bit = ((ix & 3) << 4) | ((iy & 3) << 2) | (iz & 3);
        return (word >> bit) & 1ull;
    }
};

// ---------------------------------------------------------------------------
// Ray / box slab test: reference utils_grid.cuh:10-55.
// ---------------------------------------------------------------------------
NFA_HD bool slab_test(const float o[3], const float inv[3], const float* box,
                      float near, float far, float& tmin, float& tmax)
{
    float lo = 0.f, hi = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float t_lo, t_hi;
        if (inv[a] >= 0.0f) {
            t_lo = f_mul(f_sub(box[a], o[a]), inv[a]);
            t_hi = f_mul(f_sub(box[3 + a], o[a]), inv[a]);
        } else {
            t_lo = f_mul(f_sub(box[3 + a], o[a]), inv[a]);
            t_hi = f_mul(f_sub(box[a], o[a]), inv[a]);
        }
        if (a == 0) {
            lo = t_lo;
            hi = t_hi;
        } else {
            if (lo > t_hi || t_lo > hi) return false;
            if (t_lo > lo) lo = t_lo;
            if (t_hi < hi) hi = t_hi;
        }
    }
    if (hi <= 0.0f) return false;
    tmin = f_max(lo, near);
    tmax = f_min(hi, far);
    return true;
}

// ---------------------------------------------------------------------------
// DDA over one (ray, level, [seg_lo, seg_hi]) segment.
// Set-up: reference utils_grid.cuh:58-114; step: utils_grid.cuh:116-142.
// ---------------------------------------------------------------------------
struct Dda {
    float td[3], dl[3];
    int cur[3], st[3], ov[3];
};

NFA_HD void dda_begin(Dda& s, const float o[3], const float d[3], const float inv[3],
                      float tmin, float tmax, const float* box, const int res[3])
{
    const float eps = 1e-6f;  // reference grid.cu:95
    const float t_in = f_add(tmin, eps);
    const float t_out = f_add(tmax, -eps);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float resf = (float)res[a];
        const float extent = f_sub(box[3 + a], box[a]);
        const float voxel = f_div(extent, resf);
        const float p_in = f_fma(d[a], t_in, o[a]);
        const float p_out = f_fma(d[a], t_out, o[a]);
        const int first = i_clamp(f_trunc_i32(f_mul(f_div(f_sub(p_in, box[a]), extent), resf)), 0, res[a] - 1);
        const int last = i_clamp(f_trunc_i32(f_mul(f_div(f_sub(p_out, box[a]), extent), resf)), 0, res[a] - 1);
        const int ahead = d[a] > 0.0f ? 1 : 0;
        const float face = (float)(uint32_t)(first + ahead);
        const float cross = f_fma(inv[a], f_add(box[a], f_fma(face, voxel, -p_in)), tmin);
        if (d[a] == 0.0f) {
            s.td[a] = tmax;
            s.dl[a] = tmax;
            s.st[a] = 0;
        } else {
            s.td[a] = cross;
            const float m = f_mul(inv[a], voxel);
            s.dl[a] = d[a] > 0.0f ? m : -m;
            s.st[a] = d[a] > 0.0f ? 1 : -1;
        }
        s.cur[a] = first;
        s.ov[a] = last + s.st[a];
    }
}

// ---------------------------------------------------------------------------
// Run collection.  A run is a maximal chain of consecutive lattice samples with
// no empty cell (and no skipped gap) between them -- exactly the reference's
// `continuous` chains (grid.cu:219-245,258-259), so runs also delimit the
// interval edges traverse_grids() returns.
// ---------------------------------------------------------------------------
struct RayMarch {
    Lattice L;
    float t;           // lattice anchor; the reference's t_last is seek(t, pend)
    float pend;        // pending skip target (-inf: none)
    float open_target; // exit time of the last occupied cell not yet counted
    float run_first;
    uint32_t run_n;    // samples counted so far in the open run (> 0 <=> `continuous`)
    bool dirty;        // occupied cells seen since the last settle()
    bool ok;           // false once the lattice got stuck / guard tripped
    // results
    uint32_t n_samples;
    uint32_t n_runs;
};

NFA_HD void rm_init(RayMarch& m, const Lattice& L, float near)
{
    m.L = L;
    m.t = near;
    m.pend = -INFINITY;
    m.open_target = 0.f;
    m.run_first = 0.f;
    m.run_n = 0;
    m.dirty = false;
    m.ok = true;
    m.n_samples = 0;
    m.n_runs = 0;
}

// Apply the pending skip, then count the samples of the occupied cells seen
// since the last settle.
NFA_HD void rm_settle(RayMarch& m)
{
    if (!m.dirty || !m.ok) { m.dirty = false; return; }
    uint32_t k = 0;
    if (m.pend > -INFINITY) {
        m.ok = lat_seek(m.L, m.t, m.pend, k);
        m.pend = -INFINITY;
        if (!m.ok) { m.dirty = false; return; }
    }
    if (m.run_n == 0) m.run_first = m.t;
    k = 0;
    m.ok = lat_seek(m.L, m.t, m.open_target, k);
    m.run_n += k;
    m.dirty = false;
}

template <class Sink>
NFA_HD void rm_close_run(RayMarch& m, Sink& sink)
{
    if (m.run_n > 0) {
        sink.push(m.n_runs, m.run_first, m.run_n);
        m.n_runs += 1;
        m.n_samples += m.run_n;
        m.run_n = 0;
    }
}

// Walk one segment.  `level` selects the grid, [seg_lo, seg_hi] is already
// clipped to the ray's near/far planes (grid.cu:148-150).
template <class Sink>
NFA_HD void rm_segment(RayMarch& m, Sink& sink, const OccView& occ, OccCursor& cur,
                       const float o[3], const float d[3], const float inv[3],
                       int level, float seg_lo, float seg_hi, const float* box)
{
    // grid.cu:153-163: skip to the segment start unless we are inside a chain.
    rm_settle(m);
    if (m.run_n == 0) m.pend = f_max(m.pend, seg_lo);

    Dda s;
    dda_begin(s, o, d, inv, seg_lo, seg_hi, box, occ.g.res);
    // cells outside the grid can only be reached where the reference itself
    // reads out of bounds; stop there.
    for (int guard = 0; guard < (1 << 20) && m.ok; ++guard) {
        if ((unsigned)s.cur[0] >= (unsigned)occ.g.res[0] || (unsigned)s.cur[1] >= (unsigned)occ.g.res[1] ||
            (unsigned)s.cur[2] >= (unsigned)occ.g.res[2])
            break;
        const float tt = f_min(f_min(s.td[0], f_min(s.td[1], s.td[2])), seg_hi);  // grid.cu:185-186
        if (cur.test(occ, level, s.cur[0], s.cur[1], s.cur[2])) {
            m.open_target = tt;  // grid.cu:206-262, evaluated lazily
            m.dirty = true;
        } else {
            // grid.cu:194-205: an empty cell ends the chain and skips past its exit
            if (m.dirty) rm_settle(m);
            rm_close_run(m, sink);
            m.pend = f_max(m.pend, tt);
        }
        // utils_grid.cuh:116-142
        int a;
        if (s.td[0] < s.td[1] && s.td[0] < s.td[2]) a = 0;
        else if (s.td[1] < s.td[2]) a = 1;
        else a = 2;
        if (a == 0) { s.cur[0] += s.st[0]; s.td[0] = f_add(s.td[0], s.dl[0]); if (s.cur[0] == s.ov[0]) break; }
        else if (a == 1) { s.cur[1] += s.st[1]; s.td[1] = f_add(s.td[1], s.dl[1]); if (s.cur[1] == s.ov[1]) break; }
        else { s.cur[2] += s.st[2]; s.td[2] = f_add(s.td[2], s.dl[2]); if (s.cur[2] == s.ov[2]) break; }
    }
}

// End of ray: count what is still open; optionally report the reference's
// terminate plane (grid.cu:274-275).
template <class Sink>
NFA_HD float rm_finish(RayMarch& m, Sink& sink, bool want_terminate)
{
    rm_settle(m);
    rm_close_run(m, sink);
    if (want_terminate && m.ok && m.pend > -INFINITY) {
        uint32_t k = 0;
        m.ok = lat_seek(m.L, m.t, m.pend, k);
        m.pend = -INFINITY;
    }
    return m.t;
}

// ---------------------------------------------------------------------------
// Whole-ray drivers.
// ---------------------------------------------------------------------------

// Single grid level, box crossings computed in place (what the reference does
// with ray_aabb_intersect + torch.sort for n_grids == 1, grid.py:156-162).
template <class Sink>
NFA_HD float march_ray_single(RayMarch& m, Sink& sink, const OccView& occ,
                              const float o[3], const float d[3], float near, float far,
                              const float* box, const Lattice& L, bool want_terminate)
{
    const float inv[3] = {f_rcp(d[0]), f_rcp(d[1]), f_rcp(d[2])};
    rm_init(m, L, near);
    OccCursor cur;
    cur.reset();
    float tmin, tmax;
    if (slab_test(o, inv, box, -INFINITY, INFINITY, tmin, tmax)) {
        const float seg_lo = f_max(tmin, near);
        const float seg_hi = f_min(tmax, far);
        if (!(seg_lo >= seg_hi)) rm_segment(m, sink, occ, cur, o, d, inv, 0, seg_lo, seg_hi, box);
    }
    return rm_finish(m, sink, want_terminate);
}

// Any number of levels, crossings given sorted (grid.cu:129-150).
template <class Sink>
NFA_HD float march_ray_sorted(RayMarch& m, Sink& sink, const OccView& occ,
                              const float o[3], const float d[3], float near, float far,
                              const float* aabbs, int n_grids,
                              const float* t_sorted, const int64_t* t_indices, const uint8_t* hits,
                              const Lattice& L, bool want_terminate)
{
    const float inv[3] = {f_rcp(d[0]), f_rcp(d[1]), f_rcp(d[2])};
    rm_init(m, L, near);
    OccCursor cur;
    cur.reset();
    for (int i = 0; i < 2 * n_grids - 1 && m.ok; ++i) {
        const int64_t id = t_indices[i];
        int level = (int)(id % n_grids);
        if (!hits[level]) continue;
        if (!(id < n_grids)) {  // leaving this box: only go on if still inside the next one
            const int64_t nx = t_indices[i + 1];
            if (nx < n_grids) continue;
            level = (int)(nx % n_grids);
            if (!hits[level]) continue;
        }
        const float seg_lo = f_max(t_sorted[i], near);
        const float seg_hi = f_min(t_sorted[i + 1], far);
        if (seg_lo >= seg_hi) continue;
        rm_segment(m, sink, occ, cur, o, d, inv, level, seg_lo, seg_hi, aabbs + 6 * level);
    }
    return rm_finish(m, sink, want_terminate);
}

}  // namespace nfa
