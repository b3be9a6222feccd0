/*  plan.h — helper that cuts an output panel into cells/tiles for the chain kernel (see plan.cpp). */
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.h"

namespace dmrgx {

/* One rectangular contribution to an output panel: rows [r0,r0+nr) × cols [c0,c0+nc) in panel
   coordinates; `seg` addresses its operands at the rectangle's origin. */
struct Contribution {
    int r0, c0, nr, nc;
    dev::Segment seg;
};

/* "pointer" that is really a BYTE offset into the x (or y) vector of the launch: ordinary pointer
   arithmetic on it keeps working, the kernel adds the base when SEGF_*_X / c_in_y is set */
inline const double* xoff(long long elems) { return (const double*)(uintptr_t)(elems * 8); }
inline double* yoff(long long elems) { return (double*)(uintptr_t)(elems * 8); }

inline dev::Segment make_seg(int type) {
    dev::Segment s;
    std::memset(&s, 0, sizeof s);
    s.type = type;
    s.coef = 1.0;
    return s;
}

void emit_cells(Plan& plan, double* C, bool c_in_y, long long ldc, int R, int Ncols, const std::vector<Contribution>& contribs,
                bool cover_all);

/* Contribution of `coef * tile` (any format) ADDED into an output rectangle of the same shape */
Contribution add_tile_contribution(const Tile& t, int r0, int c0, double coef);

}  // namespace dmrgx
