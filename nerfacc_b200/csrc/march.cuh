// march.cuh -- per-ray occupancy-grid traversal, constant-step fast path.
//
// Restates the semantics of the reference's traverse_grids_kernel
// (/root/reference/nerfacc/cuda/csrc/grid.cu:68-282 with the helpers in
// include/utils_grid.cuh:10-142) for cone_angle == 0 and step_size > 0, split
// into two phases so that a warp of unrelated rays stays convergent:
//
//  phase 1 (walk_*)  the DDA walks cells exactly like the reference (same f32
//                    evaluation order and FMA contraction as the reference's
//                    sm_100a SASS) but does no marching at all:
//                    it only records *stretches* -- maximal chains of occupied
//                    cells -- as (pend, open) pairs in the time domain:
//                    `pend` = where the reference's skip loops would have moved
//                    t_last before the stretch, `open` = exit time of its last cell.
//  phase 2 (lat_*)   turns stretches into runs (t_first, n) of consecutive lattice
//                    samples with two closed-form seeks each (lattice.cuh).  All
//                    lanes do this at the same time, so the expensive integer
//                    code is not serialised by divergence.
//
// The per-sample arrays are produced later by the expand kernel from the runs.
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
// Phase 1: the walk.  Events of the reference kernel, per ray:
//   SEG(lo)   segment start: `if (!continuous)` skip to lo        grid.cu:153-163
//   EMPTY(tt) empty cell: skip past its exit tt, continuous=false grid.cu:194-205
//   OCC(tt)   occupied cell: emit samples while mid < tt          grid.cu:206-262
// A stretch is a maximal event sequence without EMPTY.  `continuous` is false at
// its start and turns true at its first emitted sample, so inside a stretch a
// later SEG only skips if nothing was emitted yet: such a SEG splits the stretch
// into a second, `joined` descriptor whose pend is applied conditionally.
// ---------------------------------------------------------------------------
struct Walk {
    // ray
    float o[3], d[3], inv[3];
    float near, far;
    // segment iteration
    int seg_i;       // next crossing index to look at
    int level;
    float seg_hi;
    bool in_seg;
    Dda s;
    OccCursor cur;
    // stretch under construction
    bool open;       // inside a stretch (no EMPTY since it began)
    bool has_occ;    // the current descriptor has seen an occupied cell
    bool joined;     // the current descriptor continues the previous one (SEG inside a stretch)
    float d_pend, d_open;
    float pend_acc;  // skip target accumulated while no stretch is open
    bool done;
};

NFA_HD void walk_init(Walk& w, const float o[3], const float d[3], float near, float far)
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        w.o[a] = o[a];
        w.d[a] = d[a];
        w.inv[a] = f_rcp(d[a]);
    }
    w.near = near;
    w.far = far;
    w.seg_i = 0;
    w.level = 0;
    w.seg_hi = 0.f;
    w.in_seg = false;
    w.cur.reset();
    w.open = false;
    w.has_occ = false;
    w.joined = false;
    w.d_pend = -INFINITY;
    w.d_open = 0.f;
    w.pend_acc = -INFINITY;
    w.done = false;
}

// Crossing source for a single grid level: what ray_aabb_intersect + torch.sort
// produce for n_grids == 1 (reference grid.py:156-162), computed in place.
struct SingleBox {
    const float* box;
    NFA_HD int n_grids() const { return 1; }
    NFA_HD const float* aabb(int) const { return box; }
    // next valid segment at or after index i (i is advanced past it); false when exhausted
    NFA_HD bool next(const Walk& w, int& i, int& level, float& lo, float& hi) const
    {
        if (i > 0) return false;
        i = 1;
        float tmin, tmax;
        if (!slab_test(w.o, w.inv, box, -INFINITY, INFINITY, tmin, tmax)) return false;
        lo = f_max(tmin, w.near);
        hi = f_min(tmax, w.far);
        level = 0;
        return !(lo >= hi);
    }
};

// Crossings given sorted (reference grid.cu:129-150).
struct SortedBoxes {
    const float* aabbs;
    int G;
    const float* t_sorted;     // this ray's [2G]
    const int64_t* t_indices;  // this ray's [2G]
    const uint8_t* hits;       // this ray's [G]
    NFA_HD int n_grids() const { return G; }
    NFA_HD const float* aabb(int level) const { return aabbs + 6 * level; }
    NFA_HD bool next(const Walk& w, int& i, int& level, float& lo, float& hi) const
    {
        for (; i < 2 * G - 1; ++i) {
            const int64_t id = t_indices[i];
            int lv = (int)(id % G);
            if (!hits[lv]) continue;
            if (!(id < G)) {  // leaving this box: go on only if still inside the next one
                const int64_t nx = t_indices[i + 1];
                if (nx < G) continue;
                lv = (int)(nx % G);
                if (!hits[lv]) continue;
            }
            lo = f_max(t_sorted[i], w.near);
            hi = f_min(t_sorted[i + 1], w.far);
            if (lo >= hi) continue;
            level = lv;
            ++i;
            return true;
        }
        return false;
    }
};

// One step of the walk: either opens the next segment or processes one cell.
// Descriptors are appended through `buf.put(j, pend, open, joined)`; the caller
// stops calling when its buffer is full and resumes after flushing it.
template <class Boxes, class Buf>
NFA_HD void walk_step(Walk& w, const Boxes& boxes, const OccView& occ, Buf& buf, int& n_desc)
{
    if (!w.in_seg) {
        int level;
        float lo, hi;
        if (!boxes.next(w, w.seg_i, level, lo, hi)) {
            // a descriptor without occupied cells (open = -inf) still carries its conditional skip
            if (w.open) buf.put(n_desc++, w.d_pend, w.has_occ ? w.d_open : -INFINITY, w.joined);
            w.open = false;
            w.done = true;
            return;
        }
        // SEG(lo)
        if (w.open) {
            buf.put(n_desc++, w.d_pend, w.has_occ ? w.d_open : -INFINITY, w.joined);
            w.d_pend = lo;
            w.joined = true;
            w.has_occ = false;
        } else {
            w.pend_acc = f_max(w.pend_acc, lo);
        }
        w.level = level;
        w.seg_hi = hi;
        dda_begin(w.s, w.o, w.d, w.inv, lo, hi, boxes.aabb(level), occ.g.res);
        w.in_seg = true;
        return;
    }
    Dda& s = w.s;
    // cells outside the grid can only be reached where the reference itself reads out of bounds
    if ((unsigned)s.cur[0] >= (unsigned)occ.g.res[0] || (unsigned)s.cur[1] >= (unsigned)occ.g.res[1] ||
        (unsigned)s.cur[2] >= (unsigned)occ.g.res[2]) {
        w.in_seg = false;
        return;
    }
    const float tt = f_min(f_min(s.td[0], f_min(s.td[1], s.td[2])), w.seg_hi);  // grid.cu:185-186
    if (w.cur.test(occ, w.level, s.cur[0], s.cur[1], s.cur[2])) {
        // OCC(tt)
        if (!w.open) {
            w.open = true;
            w.joined = false;
            w.d_pend = w.pend_acc;
            w.pend_acc = -INFINITY;
        }
        w.d_open = tt;
        w.has_occ = true;
    } else {
        // EMPTY(tt)
        if (w.open) {
            buf.put(n_desc++, w.d_pend, w.has_occ ? w.d_open : -INFINITY, w.joined);
            w.open = false;
            w.has_occ = false;
        }
        w.pend_acc = f_max(w.pend_acc, tt);
    }
    // utils_grid.cuh:116-142: x only if strictly smallest, else y if strictly below z, else z
    int a;
    if (s.td[0] < s.td[1] && s.td[0] < s.td[2]) a = 0;
    else if (s.td[1] < s.td[2]) a = 1;
    else a = 2;
    bool leave;
    if (a == 0) { s.cur[0] += s.st[0]; s.td[0] = f_add(s.td[0], s.dl[0]); leave = s.cur[0] == s.ov[0]; }
    else if (a == 1) { s.cur[1] += s.st[1]; s.td[1] = f_add(s.td[1], s.dl[1]); leave = s.cur[1] == s.ov[1]; }
    else { s.cur[2] += s.st[2]; s.td[2] = f_add(s.td[2], s.dl[2]); leave = s.cur[2] == s.ov[2]; }
    if (leave) w.in_seg = false;
}

// Skip target still pending at the end of the ray (for the terminate plane, grid.cu:274-275).
NFA_HD float walk_tail_pend(const Walk& w)
{
    return w.open ? -INFINITY : w.pend_acc;
}

// ---------------------------------------------------------------------------
// Phase 2: stretches -> runs on the lattice.  A run is a maximal chain of
// consecutive lattice samples, exactly the reference's `continuous` chains
// (grid.cu:219-245,258-259), so runs also delimit traverse_grids()' interval edges.
// ---------------------------------------------------------------------------
struct LatState {
    Lattice L;
    float t;          // lattice anchor (the reference's t_last once pending skips are applied)
    float run_first;
    uint32_t run_n;   // samples in the open run (> 0 <=> the reference's `continuous`)
    uint32_t n_samples, n_runs;
    bool ok;          // false once the lattice got stuck (the reference would not terminate)
};

struct RunOut {
    bool valid;
    float t_first;
    uint32_t n;
    uint32_t sample_off;  // samples of this ray before the run
    uint32_t run_idx;     // runs of this ray before the run
};

NFA_HD void lat_init(LatState& m, const Lattice& L, float near)
{
    m.L = L;
    m.t = near;
    m.run_first = 0.f;
    m.run_n = 0;
    m.n_samples = 0;
    m.n_runs = 0;
    m.ok = true;
}

NFA_HD void lat_close(LatState& m, RunOut& out)
{
    out.valid = m.run_n > 0;
    if (out.valid) {
        out.t_first = m.run_first;
        out.n = m.run_n;
        out.sample_off = m.n_samples;
        out.run_idx = m.n_runs;
        m.n_samples += m.run_n;
        m.n_runs += 1;
        m.run_n = 0;
    }
}

// Consume one descriptor; `out` reports the run that this descriptor closed, if any.
NFA_HD void lat_consume(LatState& m, float pend, float open, bool joined, RunOut& out)
{
    out.valid = false;
    if (!joined) lat_close(m, out);  // an EMPTY cell separated this stretch from the previous one
    if (!m.ok) return;
    uint32_t k = 0;
    if (m.run_n == 0) {
        m.ok = lat_seek(m.L, m.t, pend, k);
        if (!m.ok) return;
        m.run_first = m.t;
    }
    k = 0;
    m.ok = lat_seek(m.L, m.t, open, k);
    m.run_n += k;
}

// End of ray: close the open run and (optionally) apply the trailing skip.
NFA_HD float lat_finish(LatState& m, float tail_pend, bool want_terminate, RunOut& out)
{
    lat_close(m, out);
    if (want_terminate && m.ok && tail_pend > -INFINITY) {
        uint32_t k = 0;
        m.ok = lat_seek(m.L, m.t, tail_pend, k);
    }
    return m.t;
}

}  // namespace nfa
