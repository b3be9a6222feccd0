/*  truncate.cpp — GetTruncation and RotateOperators on the device.
 *
 *  include/DMRGBlockContainer.hpp:1656-1959 (GetTruncation), :1962-2003 (EigRDM_BlockDiag),
 *  :2006-2057 (FillRotation_BlockDiag), :82-94 (Eigen_t sort keys);  src/DMRGBlock.cpp:677-823
 *  (RotateOperators: O' = RotMatT · O · RotMatTᴴ for every Sz_i, Sp_i and H).
 *
 *  The reference gathers ψ to rank 0 and works serially; here ψ never leaves HBM:
 *    ρ_L = X_p X_pᵀ, ρ_R = X_pᵀ X_p for all sector pairs   -> one chain launch (FP64 DMMA)
 *    per-block eigendecomposition                           -> cuSOLVER syevd (LAPACK on the reference side)
 *    global top-m selection                                 -> host, on the eigenvalue lists only, with the
 *                                                              reference's two stable sorts (bit-exact counts)
 *    O' = U_I · O[I,J] · U_Jᵀ for all operators and blocks   -> two chain launches
 */
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"
#include "plan.h"

namespace dmrgx {

namespace {
struct Eigen_t { double eigval; int seqIdx, epsIdx, blkIdx; };
}

/* one side (left or right) of GetTruncation, in two phases so that the eigendecompositions of BOTH sides go out as
   one batch over the solver lanes */
struct SideJob {
    bool left;
    std::vector<long long> roff, woff;
    std::vector<int> dim, blkidx, owner; /* owner: rank that builds and diagonalises the block */
    long long wtot = 0;
    BufRef rho, dw;
};

static void side_layout(const Kron* kron, bool left, SideJob& J) {
    Ctx* ctx = kron->ctx;
    const Sectors &SL = kron->L->sec, &SR = kron->R->sec;
    const int np = (int)kron->pairs.size();
    J.left = left;
    /* ---- ρ blocks, one per sector pair, in KronBlocks order (:1709-1775) ---- */
    J.roff.assign(np + 1, 0); J.woff.assign(np + 1, 0);
    J.dim.resize(np); J.blkidx.resize(np);
    for (int p = 0; p < np; ++p) {
        J.dim[p] = left ? SL.size[kron->pairs[p].il] : SR.size[kron->pairs[p].ir];
        J.blkidx[p] = left ? kron->pairs[p].il : kron->pairs[p].ir;
        J.roff[p + 1] = J.roff[p] + (long long)J.dim[p] * J.dim[p];
        J.woff[p + 1] = J.woff[p] + J.dim[p];
    }
    J.wtot = J.woff[np];
    J.rho = std::make_shared<DevBuf>(ctx, std::max<long long>(1, J.roff.back()) * 8);
    J.dw = std::make_shared<DevBuf>(ctx, std::max<long long>(1, J.wtot) * 8);
    J.owner.assign(np, 0);
}

static void side_build_rho(const Kron* kron, const double* d_psi, SideJob& J) {
    Ctx* ctx = kron->ctx;
    const Sectors &SL = kron->L->sec, &SR = kron->R->sec;
    const int np = (int)kron->pairs.size();
    const bool left = J.left;
    Plan plan;
    for (int p = 0; p < np; ++p) {
        const int nL = SL.size[kron->pairs[p].il], nR = SR.size[kron->pairs[p].ir];
        const int n = J.dim[p];
        if (n == 0 || J.owner[p] != ctx->rank) continue;
        Contribution c;
        c.r0 = 0; c.c0 = 0; c.nr = n; c.nc = n;
        c.seg = make_seg(dev::SEG_GEMM);
        c.seg.flags = dev::SEGF_A_X | dev::SEGF_B_X;
        c.seg.A = xoff(kron->off[p]);
        c.seg.B = xoff(kron->off[p]);
        if (left) { /* rdmd_L = Psi · PsiT, Psi row-major nL×nR (:1731-1733) */
            c.seg.lda_m = nR; c.seg.lda_k = 1; c.seg.ldb_n = nR; c.seg.ldb_k = 1; c.seg.K = nR;
        } else {    /* rdmd_R = PsiT · Psi (:1734) */
            c.seg.lda_m = 1; c.seg.lda_k = nR; c.seg.ldb_n = 1; c.seg.ldb_k = nR; c.seg.K = nL;
        }
        if (c.seg.K == 0) continue;
        std::vector<Contribution> cs = {c};
        emit_cells(plan, J.rho->as<double>() + J.roff[p], false, n, n, n, cs, true);
    }
    plan.upload(ctx);
    plan.run(ctx, d_psi, nullptr);
    dev::sync(ctx->st); /* the plan's device lists die with this scope */
}

static XForm* side_select(const Kron* kron, long long mstates, const SideJob& J) {
    Ctx* ctx = kron->ctx;
    dev::Stream* st = ctx->st;
    const bool left = J.left;
    const Block* blk = left ? kron->L : kron->R;
    const Sectors& S = blk->sec;
    const int np = (int)kron->pairs.size();
    const std::vector<long long>&roff = J.roff, &woff = J.woff;
    const std::vector<int>&dim = J.dim, &blkidx = J.blkidx;
    const long long wtot = J.wtot;
    const BufRef &rho = J.rho, &dw = J.dw;
    std::vector<double> w(std::max<long long>(1, wtot));
    dev::d2h(st, w.data(), dw->p, (size_t)wtot * 8);
    dev::sync(st);
    /* ---- selection: exactly the reference's ordering rules ---- */
    std::vector<Eigen_t> eigen;
    for (int p = 0; p < np; ++p)
        for (int kx = 0; kx < dim[p]; ++kx) /* EPS_LARGEST_REAL: descending inside a block */
            eigen.push_back({w[woff[p] + (dim[p] - 1 - kx)], p, kx, blkidx[p]});
    std::unique_ptr<XForm> xf(new XForm());
    xf->ctx = ctx;
    xf->nstates_old = S.nstates();
    for (const Eigen_t& e : eigen) { xf->spec_eig.push_back(e.eigval); xf->spec_blk.push_back(e.blkIdx); }
    std::stable_sort(eigen.begin(), eigen.end(), [](const Eigen_t& a, const Eigen_t& b) { return a.eigval > b.eigval; }); /* :1795 */
    const long long m = std::min<long long>(mstates, (long long)eigen.size());                                            /* :1819 */
    eigen.resize((size_t)m);
    std::stable_sort(eigen.begin(), eigen.end(), [](const Eigen_t& a, const Eigen_t& b) { return a.blkIdx < b.blkIdx; }); /* :1850 */
    xf->trunc_err = 1.0;
    for (const Eigen_t& e : eigen) xf->trunc_err -= (e.eigval > 0) * e.eigval; /* :1872-1875 */
    /* new sector list: blocks that kept at least one state (:1879-1883) */
    std::map<int, int> kept;      /* blkIdx -> count */
    std::map<int, int> seq_of;    /* blkIdx -> seqIdx */
    for (const Eigen_t& e : eigen) {
        kept[e.blkIdx] += 1;
        auto f = seq_of.find(e.blkIdx);
        if (f == seq_of.end()) seq_of[e.blkIdx] = e.seqIdx;
        else if (f->second != e.seqIdx) throw Err(ERR_SUP, "a block sector appears in more than one sector pair (multi-sector targets are not supported in the truncation)");
    }
    std::vector<double> qn;
    std::vector<long long> sz;
    for (auto& kv : kept) {
        qn.push_back(S.qn[kv.first]); sz.push_back(kv.second); xf->old_sector.push_back(kv.first);
        xf->old_off.push_back(S.off[kv.first]); xf->old_size.push_back(S.size[kv.first]);
    }
    if (qn.empty()) throw Err(ERR_GENERIC, "truncation kept no states");
    xf->newsec.init(qn, sz);
    /* ---- rows of RotMatT: the m_I leading eigenvectors of each kept block, descending (:2006-2057) ---- */
    for (size_t k = 0; k < xf->old_sector.size(); ++k) {
        const int b = xf->old_sector[k];
        const int p = seq_of[b];
        const int n = dim[p], mI = kept[b];
        BufRef U = std::make_shared<DevBuf>(ctx, (size_t)mI * n * 8);
        dev::gather_rows_reversed(st, rho->as<double>() + roff[p], n, mI, U->as<double>());
        xf->U.push_back(U);
    }
    dev::sync(st);
    return xf.release();
}

void truncate(const Kron* kron, const double* d_psi, long long mstates, XForm** L, XForm** R) {
    Trace tr(kron->ctx, "truncate");
    Ctx* ctx = kron->ctx;
    SideJob JL, JR;
    side_layout(kron, true, JL);
    side_layout(kron, false, JR);
    /* The reference gathers psi to rank 0 and diagonalises every block there, serially (:1675-1775).  Here psi is
       already complete on every rank; the blocks of both sides are dealt to the ranks largest-first (n^3 cost, greedy
       least-loaded, the same table on every rank), each rank builds and diagonalises its own, and the eigenvectors and
       eigenvalues are then broadcast from their owners in one NCCL group. */
    if (ctx->world > 1) {
        struct Ref { SideJob* J; int p; double cost; };
        std::vector<Ref> refs;
        for (SideJob* J : {&JL, &JR})
            for (size_t p = 0; p < J->dim.size(); ++p)
                if (J->dim[p] > 0) refs.push_back({J, (int)p, std::pow((double)J->dim[p], 3.0)});
        std::stable_sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) { return a.cost > b.cost; });
        std::vector<double> load(ctx->world, 0.0);
        for (const Ref& r : refs) {
            int best = 0;
            for (int k = 1; k < ctx->world; ++k) if (load[k] < load[best]) best = k;
            r.J->owner[r.p] = best;
            load[best] += r.cost + 1e6; /* per-call latency of a small eigensolve */
        }
    }
    side_build_rho(kron, d_psi, JL);
    side_build_rho(kron, d_psi, JR);
    tr.mark("rho");
    /* ---- full spectrum of every block of both sides (EPSLAPACK, all n pairs, :1976-1994) ---- */
    std::vector<int> n;
    std::vector<double*> A, W;
    for (SideJob* J : {&JL, &JR})
        for (size_t p = 0; p < J->dim.size(); ++p) {
            if (J->dim[p] == 0 || J->owner[p] != ctx->rank) continue;
            n.push_back(J->dim[p]);
            A.push_back(J->rho->as<double>() + J->roff[p]);
            W.push_back(J->dw->as<double>() + J->woff[p]);
        }
    const int e = dev::syevd_batch(ctx->st, (int)n.size(), n.data(), A.data(), W.data());
    if (e) throw Err(ERR_GENERIC, std::string("eigendecomposition of a reduced density matrix block failed: ") + dev::last_error());
    if (ctx->world > 1) {
        std::vector<double*> ptr; std::vector<long long> cnt; std::vector<int> root;
        for (SideJob* J : {&JL, &JR})
            for (size_t p = 0; p < J->dim.size(); ++p) {
                if (J->dim[p] == 0) continue;
                ptr.push_back(J->rho->as<double>() + J->roff[p]); cnt.push_back((long long)J->dim[p] * J->dim[p]); root.push_back(J->owner[p]);
                ptr.push_back(J->dw->as<double>() + J->woff[p]); cnt.push_back(J->dim[p]); root.push_back(J->owner[p]);
            }
        dev::bcast_batch(ctx->st, (int)ptr.size(), ptr.data(), cnt.data(), root.data());
    }
    tr.mark("syevd_batch");
    std::unique_ptr<XForm> l(side_select(kron, mstates, JL));
    std::unique_ptr<XForm> r(side_select(kron, mstates, JR));
    tr.mark("select");
    *L = l.release();
    *R = r.release();
}

/* src/DMRGBlock.cpp:677-823 */
Block* rotate(const Block* enl, const XForm* xf) {
    Ctx* ctx = enl->ctx;
    Trace tr(ctx, "rotate");
    const Sectors& SO = enl->sec;
    if (xf->nstates_old != SO.nstates()) throw Err(ERR_GENERIC, "RotMatT_in incorrect number of cols.");
    const Sectors& SN = xf->newsec;
    const int nn = SN.nsec();
    std::vector<int> new_of_old(SO.nsec(), -1);
    for (int k = 0; k < nn; ++k) new_of_old[xf->old_sector[k]] = k;
    std::vector<double> qn = SN.qn;
    std::vector<long long> sz(SN.size.begin(), SN.size.end());
    std::unique_ptr<Block> out(block_from_csr_begin(ctx, enl->nsites, qn, sz));

    /* all operators in one pass: Sp_i, Sz_i for every site, then H (:761-773) */
    struct Job { const Operator* src; Operator* dst; };
    std::vector<Job> jobs;
    for (int i = 0; i < enl->nsites; ++i) { jobs.push_back({&enl->Sp[i], &out->Sp[i]}); jobs.push_back({&enl->Sz[i], &out->Sz[i]}); }
    jobs.push_back({&enl->H, &out->H});
    /* size the T = O[I,J]·U_Jᵀ workspace and the output panels */
    long long ttot = 0, otot = 0;
    struct Blk { int job, I, J, Ip, Jp; long long toff, ooff; };
    std::vector<Blk> blks;
    /* Multi-GPU: the operators are dealt to the ranks in contiguous runs (every Sz_i / Sp_i costs the same), so each rank's
       output is ONE contiguous range of the output buffer and the exchange is a single in-place all-gather; the reference's
       counterpart is the -rot_nsubcomm split of src/DMRGBlock.cpp:700-760. */
    const int world = ctx->world, rank = ctx->rank;
    const size_t njobs = jobs.size();
    auto owner_of = [&](size_t j) { return (int)((j * (size_t)world) / njobs); };
    std::vector<long long> job_o0(jobs.size() + 1, 0);
    for (size_t j = 0; j < jobs.size(); ++j) {
        const Operator& O = *jobs[j].src;
        job_o0[j] = otot;
        for (int Ip = 0; Ip < nn; ++Ip) {
            const int I = xf->old_sector[Ip], J = I + O.shift;
            if (J < 0 || J >= SO.nsec()) continue;
            const int Jp = new_of_old[J];
            if (Jp < 0 || O.tiles[I].empty()) continue;
            blks.push_back({(int)j, I, J, Ip, Jp, ttot, otot});
            if (owner_of(j) == rank) ttot += (long long)SO.size[I] * SN.size[Jp];
            otot += (long long)SN.size[Ip] * SN.size[Jp];
        }
    }
    job_o0[jobs.size()] = otot;
    tr.mark("setup");
    BufRef tbuf = std::make_shared<DevBuf>(ctx, std::max<long long>(1, ttot) * 8);
    BufRef obuf = std::make_shared<DevBuf>(ctx, std::max<long long>(1, otot) * 8);
    tr.mark("alloc");
    Plan p1, p2;
    for (const Blk& b : blks) {
        const Operator& O = *jobs[b.job].src;
        const int nI = SO.size[b.I], nJ = SO.size[b.J], mI = SN.size[b.Ip], mJ = SN.size[b.Jp];
        const double* UJ = xf->U[b.Jp]->as<double>(); /* mJ × nJ */
        const double* UI = xf->U[b.Ip]->as<double>(); /* mI × nI */
        double* T = tbuf->as<double>() + b.toff;      /* nI × mJ */
        double* Oo = obuf->as<double>() + b.ooff;     /* mI × mJ */
        Tile o;
        o.fmt = T_DENSE; o.r0 = SN.off[b.Ip]; o.c0 = SN.off[b.Jp]; o.nr = mI; o.nc = mJ; o.d = Oo; o.sr = mJ; o.sc = 1; o.owner = obuf;
        jobs[b.job].dst->tiles[b.Ip].push_back(o);
        if (owner_of((size_t)b.job) != rank) continue;
        std::vector<Contribution> cs;
        for (const Tile& t : O.tiles[b.I]) {
            const int ra0 = t.r0 - SO.off[b.I], ca0 = t.c0 - SO.off[b.J];
            Contribution c;
            c.r0 = ra0; c.c0 = 0; c.nr = t.nr; c.nc = mJ;
            if (t.fmt == T_DENSE) {
                c.seg = make_seg(dev::SEG_GEMM);
                c.seg.A = t.d; c.seg.lda_m = t.sr; c.seg.lda_k = t.sc; c.seg.K = t.nc;
                c.seg.B = UJ + ca0; c.seg.ldb_n = nJ; c.seg.ldb_k = 1;
            } else if (t.fmt == T_EYE) { /* T[ra0+i, n] = scale · U_J[n, ca0+i] */
                c.seg = make_seg(dev::SEG_AXPY);
                c.seg.A = UJ + ca0; c.seg.lda_m = 1; c.seg.lda_k = nJ; c.seg.coef = t.scale;
            } else {
                c.seg = make_seg(dev::SEG_CSRA);
                c.seg.rowptr = t.rowptr; c.seg.colidx = t.col; c.seg.B = t.val;
                c.seg.A = UJ + ca0; c.seg.ldb_k = 1; c.seg.ldb_n = nJ;
            }
            cs.push_back(c);
        }
        emit_cells(p1, T, false, mJ, nI, mJ, cs, true);
        Contribution c;
        c.r0 = 0; c.c0 = 0; c.nr = mI; c.nc = mJ;
        c.seg = make_seg(dev::SEG_GEMM);
        c.seg.A = UI; c.seg.lda_m = nI; c.seg.lda_k = 1; c.seg.K = nI;
        c.seg.B = T; c.seg.ldb_k = mJ; c.seg.ldb_n = 1;
        std::vector<Contribution> c2 = {c};
        emit_cells(p2, Oo, false, mJ, mI, mJ, c2, true);
    }
    for (Job& j : jobs) { j.dst->shift = j.src->shift; j.dst->present = true; }
    tr.mark("plan");
    p1.upload(ctx);
    p2.upload(ctx);
    tr.mark("upload");
    p1.run(ctx);
    tr.mark("run1");
    p2.run(ctx);
    tr.mark("run2");
    if (world > 1) {
        std::vector<long long> cuts(world + 1, otot);
        cuts[0] = 0;
        for (int r = 1; r < world; ++r) { /* first output element of the first operator owned by a rank >= r */
            size_t j = 0;
            while (j < njobs && owner_of(j) < r) ++j;
            cuts[r] = job_o0[j];
        }
        dev::allgatherv(ctx->st, obuf->as<double>(), cuts.data());
        tr.mark("bcast");
    }
    /* Sm' = (Sp')ᵀ as views (the reference rebuilds Sm on demand, src/DMRGBlock.cpp:623-636) */
    for (int i = 0; i < out->nsites; ++i) {
        Operator& sm = out->Sm[i];
        const Operator& sp = out->Sp[i];
        sm.shift = -1; sm.present = true;
        sm.tiles.assign(nn, {});
        for (int I = 0; I < nn; ++I)
            for (const Tile& t : sp.tiles[I]) {
                Tile u = t;
                u.r0 = t.c0; u.c0 = t.r0; u.nr = t.nc; u.nc = t.nr; u.sr = t.sc; u.sc = t.sr;
                sm.tiles[I + 1].push_back(u);
            }
    }
    dev::sync(ctx->st); /* plans and the T workspace die with this scope */
    block_check(out.get());
    return out.release();
}

}  // namespace dmrgx
