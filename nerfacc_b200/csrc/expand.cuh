// expand.cuh -- turning a run (t_first, n) back into its n lattice samples.
//
// A run is cut into binade pieces (lattice.cuh): inside a piece the j-th sample
// start is an integer function of j, so the lanes of a warp (or a host loop)
// can produce samples independently and still match the reference's serial
// `t_next = t_last + dt` chain (reference grid.cu:215,259) bit for bit.
#pragma once

#include "lattice.cuh"

namespace nfa {

struct RunIter {
    float t;        // start of the next sample
    uint32_t left;  // samples still to produce
};

// Describe the next piece and advance.  Samples j = 0 .. count-1 of the piece
// start at piece_start(p, j) and end at start + dt (a real add).
NFA_HD uint32_t run_next_piece(const Lattice& L, RunIter& it, LatPiece& p)
{
    p = lat_piece(L, it.t);
    uint32_t c = 1u;
    if (p.regular) c = it.left < p.jmax + 1u ? it.left : p.jmax + 1u;
    const float last = lat_point(p, c - 1u);  // inc == 0 when irregular, so this is t
    it.t = f_add(last, L.dt);
    it.left -= c;
    return c;
}

NFA_HD float piece_start(const LatPiece& p, uint32_t j) { return lat_point(p, j); }

}  // namespace nfa
