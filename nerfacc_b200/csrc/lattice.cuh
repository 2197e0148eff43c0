// lattice.cuh -- exact closed-form arithmetic on the per-ray marching lattice.
//
// With a constant step dt (cone_angle == 0, step_size > 0) the reference's
// marching variable only ever changes by `t_last += dt` / `t_next = t_last + dt`
// (reference nerfacc/cuda/csrc/grid.cu:160,202,215,259).  All sample starts of
// a ray therefore lie on ONE sequence  t_{k+1} = fl(t_k + dt)  that does not
// depend on the occupancy grid.  Round-to-nearest binary32 addition of a
// constant is piecewise exactly linear: while t stays inside one binade
// [2^e, 2^(e+1)) with ulp u, t = M*u (M a 24-bit integer) and
//     fl(M*u + dt) = (M + I)*u ,   I = RNE(dt / u)
// as long as M + I < 2^24 (derivation in DESIGN.md "Lattice").  Only the step
// that leaves the binade, the steps below dt's own binade, and the
// round-half-even tie case with odd M need a real floating-point add.
//
// This lets the march kernel count the samples of an occupied stretch and find
// its first sample without stepping through them (lat_seek), and lets the
// expand kernel compute the j-th sample of a run independently per lane
// (LatPiece), both bit-identical to the reference's serial chain.
#pragma once

#include "nfa_math.cuh"

namespace nfa {

struct Lattice {
    float dt;      // step_size (> 0)
    float half;    // dt * 0.5f, exact
    uint32_t md;   // 24-bit significand of dt (implicit one set)
    uint32_t ed;   // biased exponent of dt
};

NFA_HD Lattice lat_make(float dt)
{
    Lattice L;
    L.dt = dt;
    L.half = f_mul(dt, 0.5f);
    const uint32_t b = f_bits(dt);
    L.ed = (b >> 23) & 0xffu;
    L.md = (b & 0x7fffffu) | 0x800000u;
    return L;
}

// One binade-piece of the lattice starting at t: points j = 0 .. jmax are
//   bits(j) = base + j * inc   (as IEEE bit patterns; the exponent field is in
// `base`, the significand grows by inc per step and never carries out).
// regular == false means only j = 0 (t itself) is known in closed form and the
// next point must be taken with a real add (t + dt).
struct LatPiece {
    uint32_t base;   // bit pattern of t
    uint32_t inc;    // per-step significand increment (ulps of t's binade)
    uint32_t jmax;   // largest j with a closed-form point
    bool regular;
    bool stuck;      // dt < ulp(t)/2: the reference would never advance
};

// floor(a / b) for 0 <= a < 2^24, 1 <= b < 2^24.  On the device: one approximate reciprocal (MUFU.RCP), a
// truncated product and an exact integer remainder that settles the last units -- a dozen instructions
// instead of the emulated integer division (or the out-of-line round-toward-zero float division).
NFA_HD uint32_t div_u24(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)b));
    uint32_t q = (uint32_t)__fmul_rz((float)a, r);   // off by a few units at most (both casts are exact)
    int32_t rem = (int32_t)(a - q * b);
    while (rem < 0) { --q; rem += (int32_t)b; }
    while (rem >= (int32_t)b) { ++q; rem -= (int32_t)b; }
    return q;
#else
    return a / b;
#endif
}

// quotient used only as a starting guess (the caller settles it with the exact predicate)
NFA_HD float div_estimate(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// The piece at t without its length: fills base / inc / stuck and returns true when the per-step increment is
// known (then M is t's 24-bit significand and the piece has floor((0xffffff - M) / inc) further points).
NFA_HD bool lat_piece_step(const Lattice& L, float t, LatPiece& p, uint32_t& M)
{
    const uint32_t b = f_bits(t);
    p.base = b;
    p.inc = 0;
    p.jmax = 0;
    p.regular = false;
    p.stuck = false;
    const uint32_t e = b >> 23;  // includes the sign bit: negative t => e >= 256
    if (e == 0u || e >= 255u || e < L.ed || L.ed == 0u || L.ed == 255u) return false;
    const uint32_t s = e - L.ed;
    M = (b & 0x7fffffu) | 0x800000u;
    uint32_t I;
    if (s == 0u) {
        I = L.md;
    } else if (s > 24u) {
        p.stuck = true;
        return false;
    } else {
        const uint32_t q = L.md >> s;
        const uint32_t rem = L.md & ((1u << s) - 1u);
        const uint32_t hb = 1u << (s - 1u);
        if (rem > hb) I = q + 1u;
        else if (rem < hb) I = q;
        else {
            // exact tie: round-half-even.  Once M is even it stays even and the
            // increment is the even one of {q, q+1}; with M odd take a real step.
            if (M & 1u) return false;
            I = q + (q & 1u);
        }
        if (I == 0u) {
            p.stuck = true;
            return false;
        }
    }
    p.inc = I;
    return true;
}

NFA_HD LatPiece lat_piece(const Lattice& L, float t)
{
    LatPiece p;
    uint32_t M;
    if (lat_piece_step(L, t, p, M)) {
        p.jmax = div_u24(0xffffffu - M, p.inc);
        p.regular = p.jmax > 0u;
    }
    return p;
}

NFA_HD float lat_point(const LatPiece& p, uint32_t j) { return f_from_bits(p.base + j * p.inc); }

// First lattice point t_k (k >= 0, t_0 = t) with  t_k + half >= target,
// i.e. where the reference's skip loop `while (t + dt*0.5f < target) t += dt`
// stops (grid.cu:158-162,200-204) -- equivalently, when counting from the first
// sample of an occupied stretch, k is the number of samples the reference emits
// before `t_last + dt*0.5f >= t_traverse` breaks the loop (grid.cu:214).
// `k` accumulates the number of steps taken.  Returns false if the lattice is
// stuck or the guard trips (the reference would spin forever there).
NFA_HD bool lat_seek(const Lattice& L, float& t, float target, uint32_t& k)
{
    if (target != target) return false;  // NaN target: the reference never terminates
    for (int guard = 0; guard < 1 << 16; ++guard) {
        if (f_add(t, L.half) >= target) return true;
        const LatPiece p = lat_piece(L, t);
        if (p.stuck) return false;
        if (p.regular) {
            // The whole binade lies before the target (the usual case while the lattice climbs from `near` to the
            // first occupied cell): go to its last point; the step that leaves the binade is the real add below.
            const float t_last = lat_point(p, p.jmax);
            if (!(f_add(t_last, L.half) >= target)) {
                t = t_last;
                k += p.jmax;
            } else {
                // the target is reached inside this binade: estimate the step count, then settle it with the
                // exact predicate (monotone in j), so the estimate only affects speed.
                const float inc_f = f_sub(lat_point(p, 1u), t);
                const float x = div_estimate(f_sub(f_sub(target, L.half), t), inc_f);
                uint32_t j;
                if (!(x >= 1.0f)) j = 1u;
                else if (x >= (float)p.jmax) j = p.jmax;
                else j = (uint32_t)x;
                while (j > 1u && f_add(lat_point(p, j - 1u), L.half) >= target) --j;
                while (j < p.jmax && !(f_add(lat_point(p, j), L.half) >= target)) ++j;
                t = lat_point(p, j);
                k += j;
                return true;  // point j satisfies the predicate (at the latest j == jmax does, checked above)
            }
        }
        const float tn = f_add(t, L.dt);
        if (!(tn > t)) return false;
        t = tn;
        ++k;
    }
    return false;
}

// The climb from `near` to a ray's first stretch crosses every binade between them (~10 for near = 0 and a scene a few
// units away), one piece each -- and with a uniform near plane it is the same climb for every ray.  So it is done once
// per call, on the host: pts[i] = the first lattice point >= 2^(e0 + i - 127).  A seek may then start from the entry
// one binade below its target: every earlier lattice point is below half the target, and (the table only covers
// binades above dt's) so is dt/2, hence none of them can satisfy t + dt/2 >= target.
constexpr int kLatTableMax = 40;
struct LatTable {
    float pts[kLatTableMax];
    int32_t e0;  // biased exponent of entry 0
    int32_t n;   // entries (0: no table -- per-ray near planes, a negative near plane, a stuck lattice)
};

NFA_HD void lat_table_build(const Lattice& L, float near, LatTable& T)
{
    T.n = 0;
    T.e0 = 0;
    if (!(near >= 0.0f) || L.ed == 0u || L.ed >= 255u) return;
    const uint32_t e_near = f_bits(near) >> 23;
    const uint32_t e0 = (e_near > L.ed ? e_near : L.ed) + 1u;
    T.e0 = (int32_t)e0;
    Lattice L0 = L;
    L0.half = 0.0f;  // plain "first lattice point >= target"
    float t = near;
    for (int i = 0; i < kLatTableMax && e0 + (uint32_t)i < 255u; ++i) {
        uint32_t k = 0;
        if (!lat_seek(L0, t, f_from_bits((e0 + (uint32_t)i) << 23), k)) break;
        T.pts[i] = t;
        T.n = i + 1;
    }
}

// Move t forward to the table entry below `target`'s binade, if that is ahead of t.
NFA_HD void lat_table_jump(const LatTable& T, float target, float& t)
{
    if (T.n <= 0 || !(target > 0.0f)) return;
    int idx = (int)(f_bits(target) >> 23) - 1 - T.e0;
    if (idx >= T.n) idx = T.n - 1;
    if (idx < 0) return;
    const float pt = T.pts[idx];
    if (pt > t) t = pt;
}

}  // namespace nfa
