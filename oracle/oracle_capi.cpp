/*  ORACLE — TEST INFRASTRUCTURE ONLY (see dmrg_oracle.hpp header).  Flat C API over the CPU
 *  restatement so that tests/, smoke() and bench.py's cpu_baseline leg can drive it through ctypes. */
#include "dmrg_oracle.hpp"

#include <atomic>
#include <thread>

using namespace oracle;
typedef long long ll;

namespace {
thread_local std::string g_err;
template <class F>
int guard(F&& f) {
    try { f(); return 0; }
    catch (const Error& e) { g_err = e.what(); return e.code ? e.code : 1; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
std::vector<Term> make_terms(int n, const double* a, const int* iop, const ll* isite, const int* jop, const ll* jsite) {
    std::vector<Term> t;
    for (int i = 0; i < n; ++i) t.push_back({a[i], (Op_t)iop[i], isite[i], (Op_t)jop[i], jsite[i]});
    return t;
}
CSR* pick(Block* b, int optype, ll isite) {
    if (optype == 3) return &b->H;
    if (isite < 0 || isite >= b->num_sites) return nullptr;
    if (optype == 0) return &b->SzData[isite];
    if (optype == 1) return &b->SpData[isite];
    return nullptr;
}
struct Shell {
    KronSumShell sh;
    Block* L; Block* R;
};
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

/* ---- blocks ---- */
void* orc_block_single_site(int spin_twice) { Block* b = new Block(); b->InitializeSingleSite(spin_twice); return b; }
void* orc_block_create(ll nsites, ll nsectors, const double* qn, const ll* sizes, int* err) {
    Block* b = new Block();
    int e = guard([&] { b->Initialize(nsites, std::vector<Real>(qn, qn + nsectors), std::vector<Int>(sizes, sizes + nsectors)); });
    if (err) *err = e;
    if (e) { delete b; return nullptr; }
    return b;
}
void orc_block_free(void* b) { delete (Block*)b; }
int orc_block_set_op(void* blk, int optype, ll isite, const ll* rowptr, const ll* col, const double* val) {
    Block* b = (Block*)blk;
    CSR* m = pick(b, optype, isite);
    if (!m) return ERR_ARG_OUTOFRANGE;
    const ll n = b->num_states;
    *m = CSR(n, n);
    m->rowptr.assign(rowptr, rowptr + n + 1);
    m->col.assign(col, col + rowptr[n]);
    m->val.assign(val, val + rowptr[n]);
    return 0;
}
void orc_block_info(void* blk, ll* nsites, ll* nstates, ll* nsectors) {
    Block* b = (Block*)blk;
    *nsites = b->num_sites; *nstates = b->num_states; *nsectors = b->Magnetization.NumSectors();
}
void orc_block_sectors(void* blk, double* qn, ll* sizes) {
    Block* b = (Block*)blk;
    for (ll i = 0; i < b->Magnetization.NumSectors(); ++i) { qn[i] = b->Magnetization.qn_list[i]; sizes[i] = b->Magnetization.qn_size[i]; }
}
ll orc_block_op_nnz(void* blk, int optype, ll isite) { CSR* m = pick((Block*)blk, optype, isite); return m ? m->nnz() : -1; }
int orc_block_get_op(void* blk, int optype, ll isite, ll* rowptr, ll* col, double* val) {
    CSR* m = pick((Block*)blk, optype, isite);
    if (!m) return ERR_ARG_OUTOFRANGE;
    std::copy(m->rowptr.begin(), m->rowptr.end(), rowptr);
    std::copy(m->col.begin(), m->col.end(), col);
    std::copy(m->val.begin(), m->val.end(), val);
    return 0;
}
int orc_block_check(void* blk) { return guard([&] { ((Block*)blk)->CheckOperatorBlocks(); }); }
int orc_block_check_op(void* blk, int optype_shift, int optype, ll isite) {
    Block* b = (Block*)blk;
    CSR* m = pick(b, optype, isite);
    if (!m) return ERR_ARG_OUTOFRANGE;
    return guard([&] { b->MatCheckOperatorBlocks((Op_t)optype_shift, *m); });
}

/* ---- Hamiltonian term lists: src/Hamiltonians.cpp:73-122 ---- */
ll orc_ham_terms(ll Lx, ll Ly, double J1, double Jz1, double J2, double Jz2, int bcx, int bcy, ll nsites, ll maxterms, double* a,
                 int* iop, ll* isite, int* jop, ll* jsite) {
    Hamiltonian h;
    h.Lx = Lx; h.Ly = Ly; h.J1 = J1; h.Jz1 = Jz1; h.J2 = J2; h.Jz2 = Jz2; h.BCx = (BC_t)bcx; h.BCy = (BC_t)bcy;
    std::vector<Term> t = h.H(nsites);
    for (ll i = 0; i < (ll)t.size() && i < maxterms; ++i) {
        a[i] = t[i].a; iop[i] = t[i].Iop; isite[i] = t[i].Isite; jop[i] = t[i].Jop; jsite[i] = t[i].Jsite;
    }
    return (ll)t.size();
}

/* ---- KronEye_Explicit ---- */
void* orc_kron_eye(void* L, void* R, int nterms, const double* a, const int* iop, const ll* isite, const int* jop, const ll* jsite,
                   int* err) {
    Block* out = new Block();
    int e = guard([&] { KronEye_Explicit(*(Block*)L, *(Block*)R, make_terms(nterms, a, iop, isite, jop, jsite), *out); });
    if (err) *err = e;
    if (e) { delete out; return nullptr; }
    return out;
}

/* ---- KronBlocks_t ---- */
void* orc_kron_create(void* L, void* R, int nqn, const double* qn, int* err) {
    KronBlocks_t* kb = nullptr;
    int e = guard([&] { kb = new KronBlocks_t(*(Block*)L, *(Block*)R, std::vector<Real>(qn, qn + nqn)); });
    if (err) *err = e;
    return kb;
}
void orc_kron_free(void* kb) { delete (KronBlocks_t*)kb; }
ll orc_kron_size(void* kb) { return ((KronBlocks_t*)kb)->size(); }
ll orc_kron_num_states(void* kb) { return ((KronBlocks_t*)kb)->NumStates(); }
void orc_kron_data(void* kbp, double* qn, ll* il, ll* ir, ll* size, ll* offset) {
    KronBlocks_t* kb = (KronBlocks_t*)kbp;
    for (ll i = 0; i < kb->size(); ++i) { qn[i] = kb->QN(i); il[i] = kb->LeftIdx(i); ir[i] = kb->RightIdx(i); size[i] = kb->Sizes(i); }
    for (ll i = 0; i <= kb->size(); ++i) offset[i] = kb->Offsets(i);
}
ll orc_kron_map(void* kb, ll l, ll r) { return ((KronBlocks_t*)kb)->Map(l, r); }
ll orc_kron_offsets_lr(void* kb, ll l, ll r) { return ((KronBlocks_t*)kb)->Offsets(l, r); }

/* ---- KronSumShell: KronSumConstruct (shell branch) + MatMult_KronSumShell ---- */
/* rows [r0,r1) only (r0 < 0: all rows) — the row range one MPI rank of the reference would own; y of
   orc_shell_apply* is then indexed from 0 = row r0 */
void* orc_shell_create_rows(void* kbp, int nterms, const double* a, const int* iop, const ll* isite, const int* jop, const ll* jsite,
                            ll r0, ll r1, int* err) {
    KronBlocks_t* kb = (KronBlocks_t*)kbp;
    Shell* s = new Shell();
    s->L = const_cast<Block*>(&kb->LeftBlock);
    s->R = const_cast<Block*>(&kb->RightBlock);
    int e = guard([&] { kb->KronSumConstruct(*s->L, *s->R, make_terms(nterms, a, iop, isite, jop, jsite), nullptr, &s->sh, r0, r1); });
    if (err) *err = e;
    if (e) { delete s; return nullptr; }
    return s;
}
ll orc_shell_lrows(void* s) { return ((Shell*)s)->sh.lrows; }
void* orc_shell_create(void* kbp, int nterms, const double* a, const int* iop, const ll* isite, const int* jop, const ll* jsite, int* err) {
    KronBlocks_t* kb = (KronBlocks_t*)kbp;
    Shell* s = new Shell();
    s->L = const_cast<Block*>(&kb->LeftBlock);
    s->R = const_cast<Block*>(&kb->RightBlock);
    int e = guard([&] { kb->KronSumConstruct(*s->L, *s->R, make_terms(nterms, a, iop, isite, jop, jsite), nullptr, &s->sh); });
    if (err) *err = e;
    if (e) { delete s; return nullptr; }
    return s;
}
/* KronConstruct (include/DMRGKron.hpp:309; src/DMRGKron.cpp:618-694): single term 1.0 * A ⊗ B */
void* orc_shell_create_single(void* kbp, int optype_l, ll isite_l, int optype_r, ll isite_r, int* err) {
    KronBlocks_t* kb = (KronBlocks_t*)kbp;
    Shell* s = new Shell();
    s->L = const_cast<Block*>(&kb->LeftBlock);
    s->R = const_cast<Block*>(&kb->RightBlock);
    int e = guard([&] {
        if (optype_l == OpSm && !s->L->init_Sm) s->L->CreateSm();
        if (optype_r == OpSm && !s->R->init_Sm) s->R->CreateSm();
        auto get = [](const Block& B, int op, Int i) -> const CSR* { return op == OpSp ? &B.Sp(i) : (op == OpSm ? &B.Sm(i) : &B.Sz(i)); };
        s->sh.Nrows = kb->NumStates(); s->sh.rstart = 0; s->sh.lrows = s->sh.rend = kb->NumStates();
        s->sh.Terms.push_back({1.0, (Op_t)optype_l, get(*s->L, optype_l, isite_l), (Op_t)optype_r, get(*s->R, optype_r, isite_r)});
        kb->KronSumSetUpShellTerms(s->sh);
    });
    if (err) *err = e;
    if (e) { delete s; return nullptr; }
    return s;
}
void orc_shell_free(void* s) { delete (Shell*)s; }
ll orc_shell_nterms(void* s) { return ((Shell*)s)->sh.Nterms; }
double orc_shell_fmas(void* s) { return ((Shell*)s)->sh.UnfactoredFMAs(); }
void orc_shell_apply(void* s, const double* x, double* y) { ((Shell*)s)->sh.MatMult(x, y); }
/* rows [r0,r1) split contiguously over nthreads host threads the way PreSplitOwnership
   (src/MiscTools.cpp:110-115) splits rows over MPI ranks; x is shared instead of all-gathered. */
void orc_shell_apply_rows(void* sp, const double* x, double* y, ll r0, ll r1, int nthreads) {
    Shell* s = (Shell*)sp;
    if (nthreads <= 1) { s->sh.MatMultRows(x, y, r0, r1); return; }
    std::vector<std::thread> th;
    const ll n = r1 - r0;
    for (int t = 0; t < nthreads; ++t) {
        ll a = r0 + (n / nthreads) * t + std::min<ll>(t, n % nthreads);
        ll b = a + n / nthreads + (t < n % nthreads ? 1 : 0);
        th.emplace_back([=] { s->sh.MatMultRows(x, y, a, b); });
    }
    for (auto& t : th) t.join();
}

/* ---- ground state ---- */
double orc_eigs(void* sp, double tol, ll ncv, ll max_it, double* psi, ll* nmatvec, double* resid) {
    Shell* s = (Shell*)sp;
    std::vector<Real> v;
    EigsStats st;
    double e = LanczosSmallest(s->sh.Nrows, [&](const Real* x, Real* y) { s->sh.MatMult(x, y); }, v, tol, ncv, max_it, &st);
    std::copy(v.begin(), v.end(), psi);
    if (nmatvec) *nmatvec = st.nmatvec;
    if (resid) *resid = st.resid;
    return e;
}

/* ---- truncation ---- */
void* orc_truncate(void* kbp, const double* psi, ll mstates, int left, int* err) {
    KronBlocks_t* kb = (KronBlocks_t*)kbp;
    BasisTransformation* bt = new BasisTransformation();
    int e = guard([&] { GetTruncationSide(*kb, psi, mstates, left != 0, *bt); });
    if (err) *err = e;
    if (e) { delete bt; return nullptr; }
    return bt;
}
void orc_bt_free(void* bt) { delete (BasisTransformation*)bt; }
void orc_bt_info(void* btp, ll* m, ll* nstates, ll* nsectors, double* truncerr, int* tie, ll* nspec) {
    BasisTransformation* bt = (BasisTransformation*)btp;
    *m = bt->RotMatT.nrows; *nstates = bt->RotMatT.ncols; *nsectors = bt->QN.NumSectors(); *truncerr = bt->TruncErr;
    *tie = bt->tie_at_cut; *nspec = (ll)bt->spectrum.size();
}
void orc_bt_sectors(void* btp, double* qn, ll* sizes) {
    BasisTransformation* bt = (BasisTransformation*)btp;
    for (ll i = 0; i < bt->QN.NumSectors(); ++i) { qn[i] = bt->QN.qn_list[i]; sizes[i] = bt->QN.qn_size[i]; }
}
void orc_bt_spectrum(void* btp, double* eigval, ll* blkidx) {
    BasisTransformation* bt = (BasisTransformation*)btp;
    for (size_t i = 0; i < bt->spectrum.size(); ++i) { eigval[i] = bt->spectrum[i].eigval; blkidx[i] = bt->spectrum[i].blkIdx; }
}
/* dense m×N row-major copy of RotMatT */
void orc_bt_rotmat_dense(void* btp, double* out) {
    BasisTransformation* bt = (BasisTransformation*)btp;
    std::vector<Real> d = bt->RotMatT.ToDense();
    std::copy(d.begin(), d.end(), out);
}
void* orc_rotate(void* blk_enl, void* btp, int* err) {
    Block* src = (Block*)blk_enl;
    BasisTransformation* bt = (BasisTransformation*)btp;
    Block* out = new Block();
    int e = guard([&] {
        out->Initialize(src->NumSites(), bt->QN.List(), bt->QN.Sizes());
        out->spin_twice = src->spin_twice;
        if (src->init_Sm) src->DestroySm();
        RotateOperators(*out, *src, bt->RotMatT);
    });
    if (err) *err = e;
    if (e) { delete out; return nullptr; }
    return out;
}

/* ---- whole runs: Warmup + Sweeps ---- */
void* orc_dmrg_create(ll Lx, ll Ly, double J1, double Jz1, double J2, double Jz2, int bcx, int bcy, int spin_twice, double qn_sector,
                      double eps_tol, ll eps_ncv, ll eps_max_it) {
    DMRG* d = new DMRG();
    d->Ham.Lx = Lx; d->Ham.Ly = Ly; d->Ham.J1 = J1; d->Ham.Jz1 = Jz1; d->Ham.J2 = J2; d->Ham.Jz2 = Jz2;
    d->Ham.BCx = (BC_t)bcx; d->Ham.BCy = (BC_t)bcy;
    d->spin_twice = spin_twice; d->qn_sector = qn_sector; d->eps_tol = eps_tol; d->eps_ncv = eps_ncv; d->eps_max_it = eps_max_it;
    return d;
}
void orc_dmrg_free(void* d) { delete (DMRG*)d; }
int orc_dmrg_warmup(void* dp, ll mwarmup) {
    DMRG* d = (DMRG*)dp;
    return guard([&] { d->Initialize(); d->mwarmup = mwarmup; d->Warmup(); });
}
int orc_dmrg_sweep(void* dp, ll mstates) { return guard([&] { ((DMRG*)dp)->SingleSweep(mstates); }); }
ll orc_dmrg_nsteps(void* dp) { return (ll)((DMRG*)dp)->steps.size(); }
/* ints[15], reals[3], flags[3] = {nmatvec, tie_L, tie_R} */
void orc_dmrg_step(void* dp, ll i, ll* ints, double* reals, ll* flags) {
    const StepData& s = ((DMRG*)dp)->steps[i];
    ll v[15] = {s.GlobIdx, s.LoopType, s.LoopIdx, s.StepIdx, s.NumSites_Sys, s.NumSites_Env, s.NumSites_SysEnl, s.NumSites_EnvEnl,
                s.NumStates_Sys, s.NumStates_Env, s.NumStates_SysEnl, s.NumStates_EnvEnl, s.NumStates_SysRot, s.NumStates_EnvRot,
                s.NumStates_H};
    std::copy(v, v + 15, ints);
    reals[0] = s.TruncErr_Sys; reals[1] = s.TruncErr_Env; reals[2] = s.GSEnergy;
    flags[0] = s.nmatvec; flags[1] = s.tie_L; flags[2] = s.tie_R;
}
ll orc_dmrg_step_nsectors(void* dp, ll i, int left) {
    const StepData& s = ((DMRG*)dp)->steps[i];
    return (ll)(left ? s.qn_list_L.size() : s.qn_list_R.size());
}
void orc_dmrg_step_sectors(void* dp, ll i, int left, double* qn, ll* sizes) {
    const StepData& s = ((DMRG*)dp)->steps[i];
    const auto& q = left ? s.qn_list_L : s.qn_list_R;
    const auto& z = left ? s.qn_size_L : s.qn_size_R;
    for (size_t k = 0; k < q.size(); ++k) { qn[k] = q[k]; sizes[k] = z[k]; }
}
/* borrowed pointer to sys_blocks[i] (valid until the next step) */
void* orc_dmrg_block(void* dp, ll i) {
    DMRG* d = (DMRG*)dp;
    if (i < 0 || i >= (ll)d->sys_blocks.size() || !d->sys_blocks[i].Initialized()) return nullptr;
    return &d->sys_blocks[i];
}

/* dense symmetric eigensolver (stand-in for EPSLAPACK), descending; V column k = eigenvector k */
int orc_symeig(ll n, const double* A, double* w, double* V) {
    return guard([&] {
        std::vector<Real> a(A, A + n * n), ww, vv;
        SymEigDescending(n, a, ww, vv);
        std::copy(ww.begin(), ww.end(), w);
        std::copy(vv.begin(), vv.end(), V);
    });
}

}  // extern "C"
