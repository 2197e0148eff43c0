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
    uint32_t M, c = 1u;
    if (lat_piece_step(L, it.t, p, M)) {
        const uint32_t room = 0xffffffu - M;
        if ((uint64_t)(it.left - 1u) * p.inc <= room) {
            c = it.left;  // the rest of the run lies inside this binade: no need for the piece's length
        } else {
            p.jmax = div_u24(room, p.inc);  // < it.left - 1
            p.regular = p.jmax > 0u;
            c = p.jmax + 1u;
        }
    }
    const float last = lat_point(p, c - 1u);  // c == 1 when there is no closed form: this is t
    it.t = f_add(last, L.dt);
    it.left -= c;
    return c;
}

NFA_HD float piece_start(const LatPiece& p, uint32_t j) { return lat_point(p, j); }

}  // namespace nfa
