// occ_pack.cuh -- bool occupancy grid -> 4x4x4 brick words (see march.cuh OccView).
#pragma once

#include "march.cuh"

namespace nfa {

// 64-bit word of brick (bx,by,bz) of one level.  `level_cells` points at the
// level's [rx, ry, rz] bool bytes (reference layout: estimators/occ_grid.py:73-76,
// cell id = x*ry*rz + y*rz + z, grid.cu:187-192).  Cells past the grid edge read as empty.
NFA_HD uint64_t occ_brick_word(const uint8_t* level_cells, const OccGeom& g, int bx, int by, int bz)
{
    uint64_t w = 0;
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
        const int x = bx * 4 + dx;
        if (x >= g.res[0]) break;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
            const int y = by * 4 + dy;
            if (y >= g.res[1]) break;
            const uint8_t* row = level_cells + ((int64_t)x * g.res[1] + y) * g.res[2];
#pragma unroll
            for (int dz = 0; dz < 4; ++dz) {
                const int z = bz * 4 + dz;
                if (z >= g.res[2]) break;
                if (row[z]) w |= 1ull << ((dx << 4) | (dy << 2) | dz);
            }
        }
    }
    return w;
}

// bounds layout: [n_grids][6] int32 = min brick xyz, max brick xyz (inclusive); initialised to
// (0x7f7f7f7f, 0x80808080) so that an empty level reads min > max.
constexpr int32_t kBoundsMinInit = 0x7f7f7f7f;
constexpr int32_t kBoundsMaxInit = (int32_t)0x80808080;

NFA_HD int64_t occ_coarse_words(const OccGeom& g)
{
    // class mip: 2 bits per brick, 16 bricks per word; padded to a multiple of 4 words (16 bytes) so it can
    // be moved with one bulk copy
    const int64_t n = ((int64_t)g.n_grids * g.wpl + 15) / 16;
    return (n + 3) & ~(int64_t)3;
}

}  // namespace nfa
