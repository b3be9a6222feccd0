/*  hamiltonian.cpp — Hamiltonians::J1J2XXZModel_SquareLattice term lists on the host
 *  (include/Hamiltonians.hpp:77-288, src/Hamiltonians.cpp:4-122).  Pure integer/double bookkeeping; the
 *  lists feed dmrgx_block_enlarge and dmrgx_hshell_create.  Written as an explicit bond enumeration
 *  (site -> "up" and "right" neighbours, then the two upward diagonals) that reproduces the reference's
 *  term ORDER, because the shell's term order follows Ham.H() order (src/DMRGKron.cpp:976-980).
 */
#include "common.h"

namespace dmrgx {

struct Lattice {
    long long Lx, Ly;
    int bcx, bcy; /* 0 open, 1 periodic */
    long long snake(long long ix, long long jy) const { return (ix % 2 == 0) ? ix * Ly + jy : (ix + 1) * Ly - (jy + 1); }
};

std::vector<Term> ham_terms(long long Lx, long long Ly, double J1, double Jz1, double J2, double Jz2, int bcx, int bcy, long long nsites_in) {
    const Lattice lat{Lx, Ly, bcx, bcy};
    const long long ns = nsites_in < 0 ? Lx * Ly : nsites_in;
    std::vector<Term> out;
    auto bond = [&](long long s, long long nb, double J, double Jz) {
        const long long ia = std::min(s, nb), ib = std::max(s, nb);
        if (J != 0.0) { out.push_back({J, OP_SP, ia, OP_SM, ib}); out.push_back({J, OP_SM, ia, OP_SP, ib}); }
        if (Jz != 0.0) out.push_back({Jz, OP_SZ, ia, OP_SZ, ib});
    };
    const bool nn_on = (J1 != 0.0 || Jz1 != 0.0);
    /* the reference generates NNN terms only when BOTH J2 and Jz2 are non-zero (src/Hamiltonians.cpp:101) */
    const bool nnn_on = (J2 != 0.0 && Jz2 != 0.0) && Lx > 1 && Ly > 1;
    for (long long s = 0; s < ns; ++s) {
        const long long ix = s / Ly;
        const long long jy = (ix % 2 == 0) ? (s % Ly) : (Ly - 1 - (s % Ly));
        if (nn_on) {
            if (jy < Ly - 1 || bcy) { /* above; a wrap onto itself (Ly == 1) is not a bond */
                const long long j2 = (jy + 1) % Ly, nb = lat.snake(ix, j2);
                if (nb < ns && j2 != jy) bond(s, nb, J1, Jz1);
            }
            if (ix < Lx - 1 || bcx) { /* right */
                const long long i2 = (ix + 1) % Lx, nb = lat.snake(i2, jy);
                if (nb < ns && i2 != ix) bond(s, nb, J1, Jz1);
            }
        }
        if (nnn_on) {
            const bool up = (jy < Ly - 1 || bcy);
            if (up && (ix >= 1 || bcx)) { /* upper-left */
                const long long nb = lat.snake((ix + Lx - 1) % Lx, (jy + 1) % Ly);
                if (nb < ns) bond(s, nb, J2, Jz2);
            }
            if (up && (ix < Lx - 1 || bcx)) { /* upper-right */
                const long long nb = lat.snake((ix + 1) % Lx, (jy + 1) % Ly);
                if (nb < ns) bond(s, nb, J2, Jz2);
            }
        }
    }
    return out;
}

}  // namespace dmrgx
