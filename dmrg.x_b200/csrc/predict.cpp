/*  predict.cpp — start vector of the next step's eigen-solve from this step's ground state (wave-function transformation,
 *  S. R. White, Phys. Rev. Lett. 77, 3633 (1996)).
 *
 *  NOT part of the reference: its EPSSolve starts from SLEPc's random vector (include/DMRGBlockContainer.hpp:1484-1500, no
 *  EPSSetInitialSpace anywhere in the tree) and so does dmrgx_eigs_smallest.  This file is the opt-in extension behind
 *  dmrgx_eigs_smallest_from / -wavefunction_prediction: same eigenpair to the solver's tolerance, several times fewer H·psi.
 *
 *  A sweep step works on  [A][s1][s2][B]  with the enlarged blocks  L = A (x) s1  and  R = B (x) s2  (each block is enlarged by
 *  a site "on its right"; the environment is used mirrored, include/DMRGBlockContainer.hpp:996-1088).  When the LEFT block
 *  grows the next step works on  [A'][s2][s3][B-]  with  A' = U_A (A (x) s1)  (this step's rotation) and  B = U_B (B- (x) s3)
 *  (the rotation that created B when it was the growing block), so in sector blocks
 *      psi'[(a' s2), (b- s3)]  =  sum_{(a s1), b}  U_A[a', (a s1)] · psi[(a s1), (b s2)] · U_B[b, (b- s3)] :
 *    wave_create:  Phi_p = U_A[IL'] · X_p                    for every sector pair p = (IL, IR) whose IL is kept
 *    wave_apply :  psi'[(IL' s2) rows, IR- columns] = Phi_p[:, (b-sector, s2) columns] · U_B[b-sector]
 *  and mirrored when the RIGHT block grows.  Both are lists of dense products on the chain kernel (two launches in all); every
 *  destination rectangle has exactly one source.  The index maps are the enlarged-basis order of src/DMRGKron.cpp:459-615
 *  (pairs IL-major, stable sort by descending quantum number, equal quantum numbers merged).
 */
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"
#include "plan.h"

namespace dmrgx {

/* where the piece (block sector il) (x) (site sector ir) sits in the enlarged basis: merged sector and offset inside it */
struct EnlLayout {
    std::vector<double> qn;       /* merged sectors */
    std::vector<int> size;
    std::vector<int> sector, off; /* [il * nsite + ir] */
    int nsite = 0;
    bool ok = false;
};
static EnlLayout enl_layout(const Sectors& SB, const Sectors& SS) {
    EnlLayout E;
    E.nsite = SS.nsec();
    for (int s : SS.size) if (s != 1) return E; /* multi-state site sectors interleave the block index: not handled */
    struct KB { double qn; int il, ir, size; };
    std::vector<KB> kb;
    for (int il = 0; il < SB.nsec(); ++il)
        for (int ir = 0; ir < SS.nsec(); ++ir) kb.push_back({SB.qn[il] + SS.qn[ir], il, ir, SB.size[il]});
    std::stable_sort(kb.begin(), kb.end(), [](const KB& a, const KB& b) { return a.qn > b.qn; }); /* include/DMRGKron.hpp:147-158 */
    E.sector.assign(kb.size(), -1);
    E.off.assign(kb.size(), 0);
    double last = 0;
    for (const KB& k : kb) {
        if (E.qn.empty() || k.qn < last) { E.qn.push_back(k.qn); E.size.push_back(0); } /* src/DMRGKron.cpp:560-574 */
        last = k.qn;
        E.sector[(size_t)k.il * E.nsite + k.ir] = (int)E.qn.size() - 1;
        E.off[(size_t)k.il * E.nsite + k.ir] = E.size.back();
        E.size.back() += k.size;
    }
    E.ok = true;
    return E;
}
static bool same_sectors(const std::vector<double>& qa, const std::vector<int>& sa, const Sectors& b) {
    if ((int)qa.size() != b.nsec()) return false;
    for (int i = 0; i < b.nsec(); ++i) if (qa[(size_t)i] != b.qn[i] || sa[(size_t)i] != b.size[i]) return false;
    return true;
}

static Contribution gemm(int nr, int nc, int K, const double* A, long long lda_m, long long lda_k, const double* B, long long ldb_n, long long ldb_k) {
    Contribution c;
    c.r0 = 0; c.c0 = 0; c.nr = nr; c.nc = nc;
    c.seg = make_seg(dev::SEG_GEMM);
    c.seg.K = K;
    c.seg.A = A; c.seg.lda_m = lda_m; c.seg.lda_k = lda_k;
    c.seg.B = B; c.seg.ldb_n = ldb_n; c.seg.ldb_k = ldb_k;
    return c;
}

Wave* wave_create(const Kron* kron, const double* d_psi, const XForm* xf, bool grow_left) {
    Ctx* ctx = kron->ctx;
    const Sectors& SG = grow_left ? kron->L->sec : kron->R->sec; /* the enlarged block that xf truncates */
    const Sectors& SO = grow_left ? kron->R->sec : kron->L->sec; /* the other enlarged block */
    if (xf->nstates_old != SG.nstates()) throw Err(ERR_ARG_WRONG, "wave_create: the transformation does not belong to this side of the superblock");
    std::unique_ptr<Wave> W(new Wave());
    W->ctx = ctx;
    W->grow_left = grow_left;
    W->grown = xf->newsec;
    W->other_qn = SO.qn;
    W->other_size = SO.size;
    std::vector<int> new_of_old(SG.nsec(), -1);
    for (int k = 0; k < xf->newsec.nsec(); ++k) {
        if (xf->old_size[(size_t)k] != SG.size[xf->old_sector[(size_t)k]]) throw Err(ERR_ARG_WRONG, "wave_create: sector sizes of the transformation and the block differ");
        new_of_old[(size_t)xf->old_sector[(size_t)k]] = k;
    }
    long long tot = 0;
    for (size_t p = 0; p < kron->pairs.size(); ++p) {
        const int ig = grow_left ? kron->pairs[p].il : kron->pairs[p].ir, io = grow_left ? kron->pairs[p].ir : kron->pairs[p].il;
        const int kg = new_of_old[(size_t)ig];
        if (kg < 0) continue;
        Wave::Blk b;
        b.kg = kg; b.io = io; b.off = tot; b.src = kron->off[p];
        if (xf->newsec.size[kg] == 0 || SO.size[io] == 0 || xf->old_size[(size_t)kg] == 0) continue;
        tot += (long long)xf->newsec.size[kg] * SO.size[io];
        W->blks.push_back(b);
    }
    W->phi = std::make_shared<DevBuf>(ctx, (size_t)std::max<long long>(1, tot) * 8);
    Plan plan;
    for (const Wave::Blk& b : W->blks) {
        const int m = xf->newsec.size[b.kg], n = xf->old_size[(size_t)b.kg], no = SO.size[b.io];
        const double* U = xf->U[(size_t)b.kg]->as<double>(); /* m x n */
        const double* X = d_psi + b.src;                       /* grow_left: n x no, else no x n (row-major) */
        double* Phi = W->phi->as<double>() + b.off;
        std::vector<Contribution> cs;
        if (grow_left) { /* Phi (m x no) = U · X */
            cs.push_back(gemm(m, no, n, U, n, 1, X, 1, no));
            emit_cells(plan, Phi, false, no, m, no, cs, true);
        } else {         /* Phi (no x m) = X · U^T */
            cs.push_back(gemm(no, m, n, X, n, 1, U, n, 1));
            emit_cells(plan, Phi, false, m, no, m, cs, true);
        }
    }
    plan.upload(ctx);
    plan.run(ctx);
    dev::sync(ctx->st); /* the plan dies with this scope */
    return W.release();
}

bool wave_apply(const Wave* W, const XForm* xe, const Block* site, const Kron* kn, double* d_psi_new) {
    Ctx* ctx = W->ctx;
    const bool gl = W->grow_left;
    const Sectors& SGn = gl ? kn->L->sec : kn->R->sec; /* new enlarged block on the growing side   = grown (x) site  */
    const Sectors& SSn = gl ? kn->R->sec : kn->L->sec; /* new enlarged block on the shrinking side = the old basis of xe */
    /* 1. the shrinking block of the previous step, enlarged, must be what psi was expressed in */
    const EnlLayout Eo = enl_layout(xe->newsec, site->sec);
    if (!Eo.ok) return false;
    if (Eo.qn != W->other_qn || Eo.size != W->other_size) return false;
    /* 2. the grown block, enlarged, must be the growing side of the new superblock */
    const EnlLayout Eg = enl_layout(W->grown, site->sec);
    if (!Eg.ok || !same_sectors(Eg.qn, Eg.size, SGn)) return false;
    /* 3. the old basis of xe must be the shrinking side of the new superblock */
    if (xe->nstates_old != SSn.nstates()) return false;
    for (int b = 0; b < xe->newsec.nsec(); ++b) {
        const int os = xe->old_sector[(size_t)b];
        if (os < 0 || os >= SSn.nsec() || SSn.size[os] != xe->old_size[(size_t)b]) return false;
    }
    dev::memset0(ctx->st, d_psi_new, (size_t)kn->nstates() * 8);
    Plan plan;
    const int nsite = site->sec.nsec();
    for (const Wave::Blk& blk : W->blks) {
        const int mg = W->grown.size[blk.kg];           /* kept states of the grown sector */
        const int wo = W->other_size[(size_t)blk.io];   /* extent of Phi along the shrinking side's enlarged sector */
        const double* Phi = W->phi->as<double>() + blk.off; /* grow_left: mg x wo, else wo x mg */
        for (int b = 0; b < xe->newsec.nsec(); ++b)
            for (int s = 0; s < nsite; ++s) {
                if (Eo.sector[(size_t)b * nsite + s] != blk.io) continue;
                const int p0 = Eo.off[(size_t)b * nsite + s], nb = xe->newsec.size[b];
                const int osec = xe->old_sector[(size_t)b], nold = xe->old_size[(size_t)b];
                const int gsec = Eg.sector[(size_t)blk.kg * nsite + s], g0 = Eg.off[(size_t)blk.kg * nsite + s];
                const int pn = gl ? kn->find(gsec, osec) : kn->find(osec, gsec);
                if (pn < 0 || nb == 0 || nold == 0) continue; /* outside the target sector of the new superblock */
                const double* UE = xe->U[(size_t)b]->as<double>(); /* nb x nold */
                std::vector<Contribution> cs;
                if (gl) { /* psi'[g0 + a', :] (mg x nold) = Phi[:, p0 : p0 + nb] · UE */
                    const int ncol = SSn.size[osec];
                    cs.push_back(gemm(mg, nold, nb, Phi + p0, wo, 1, UE, 1, nold));
                    emit_cells(plan, d_psi_new + kn->off[(size_t)pn] + (long long)g0 * ncol, false, ncol, mg, nold, cs, true);
                } else {  /* psi'[:, g0 + b'] (nold x mg) = UE^T · Phi[p0 : p0 + nb, :] */
                    const int ncol = SGn.size[gsec];
                    cs.push_back(gemm(nold, mg, nb, UE, 1, nold, Phi + (long long)p0 * mg, 1, mg));
                    emit_cells(plan, d_psi_new + kn->off[(size_t)pn] + g0, false, ncol, nold, mg, cs, true);
                }
            }
    }
    plan.upload(ctx);
    plan.run(ctx);
    dev::sync(ctx->st);
    return true;
}

}  // namespace dmrgx
