/*  ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the
 *  product path (dmrg.x_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 *  cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 *  A plain C++17, dependency-free CPU restatement of the superblock-diagonalisation hot path of
 *  jnvance/DMRG.x.  Every function cites the reference file:line it follows (paths relative to
 *  /root/reference).  The reference itself cannot be compiled here (every TU includes PETSc 3.8.4 /
 *  SLEPc 3.8.3 headers, neither of which is installed).
 *
 *  Pinning status:
 *    - sector/index bookkeeping (QuantumNumbers, KronBlocks_t, KronEye_Explicit index maps):
 *      PINNED against the reference's only golden vectors, tests/UnitTests_DMRGKron.cpp:39-252
 *      (tests/golden/testkron01.json, tests/test_oracle_golden.py).
 *    - CheckOperatorBlocks semantics: PINNED against tests/UnitTests_DMRGBlock.cpp:76-131 fixtures.
 *    - matvec / Lanczos / truncation / rotation arithmetic: the reference holds NO test for these and
 *      the arithmetic of EPSSolve / EPSLAPACK / MatMatMatMult lives in SLEPc 3.8.3 / PETSc 3.8.4 (not in
 *      the tree) => "parity unpinned" by the reference; pinned instead against exact diagonalisation
 *      (scipy eigsh known answers in BASELINE.md §3, regenerated in tests/test_oracle_ed.py).
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace oracle {

typedef long long Int;   /* PetscInt */
typedef double Real;     /* PetscReal / PetscScalar (real build) */

/* PETSc error codes used by the reference (petscerror.h, PETSc 3.8) */
enum {
    ERR_ARG_OUTOFRANGE = 63, ERR_ARG_WRONG = 62, ERR_ARG_WRONGSTATE = 73, ERR_ARG_CORRUPT = 64, ERR_SUP = 56
};

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define ORACLE_THROW(code, msg) throw ::oracle::Error((code), (msg))

/* include/DMRGBlock.hpp:21-27 — the operator enum doubles as the sector shift */
typedef enum { OpSm = -1, OpSz = 0, OpSp = +1, OpEye = +2 } Op_t;
typedef enum { SideLeft = 0, SideRight = 1 } Side_t;

/* ------------------------------------------------------------------------------------------------
 *  QuantumNumbers — include/QuantumNumbers.hpp:30-239, src/QuantumNumbers.cpp:9-201
 * ---------------------------------------------------------------------------------------------- */
struct QuantumNumbers {
    bool initialized = false;
    Int num_sectors = 0, num_states = 0;
    std::vector<Real> qn_list;
    std::vector<Int> qn_size, qn_offset;

    /* src/QuantumNumbers.cpp:9-52 */
    void Initialize(const std::vector<Real>& qn_list_in, const std::vector<Int>& qn_size_in) {
        if (qn_list_in.size() == 0) ORACLE_THROW(ERR_ARG_WRONG, "Initialization error: Empty input list.");
        if (qn_list_in.size() != qn_size_in.size())
            ORACLE_THROW(ERR_ARG_WRONG, "Initialization error: Input list sizes mismatch.");
        num_sectors = (Int)qn_list_in.size();
        Real qn_prev = qn_list_in[0];
        for (Int i = 1; i < num_sectors; ++i) {
            if (qn_list_in[i] >= qn_prev) ORACLE_THROW(1, "qn_list_in must be sorted descending.");
            qn_prev = qn_list_in[i];
        }
        qn_list = qn_list_in;
        qn_size = qn_size_in;
        qn_offset.assign(num_sectors + 1, 0);
        for (Int i = 1; i < num_sectors + 1; ++i) qn_offset[i] = qn_offset[i - 1] + qn_size[i - 1];
        num_states = qn_offset.back();
        initialized = true;
    }
    Int NumSectors() const { return num_sectors; }
    Int NumStates() const { return num_states; }
    const std::vector<Real>& List() const { return qn_list; }
    const std::vector<Int>& Sizes() const { return qn_size; }
    const std::vector<Int>& Offsets() const { return qn_offset; }
    /* include/QuantumNumbers.hpp:100-141 — out-of-range lookups return -1 */
    Int Sizes(Int idx) const { return (idx < 0 || idx >= num_sectors) ? -1 : qn_size[idx]; }
    Int Offsets(Int idx) const { return (idx < 0 || idx >= num_sectors) ? -1 : qn_offset[idx]; }
    Real List(Int idx) const { return qn_list.at(idx); }

    /* src/QuantumNumbers.cpp:72-96 (+ the ...Start variant, include/QuantumNumbers.hpp:159-180) */
    Int OpBlockToGlobalRangeStart(Int BlockIdx, Int BlockShift, bool& flg) const {
        if (BlockIdx < 0 || BlockIdx >= num_sectors) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "BlockIdx out of bounds");
        Int BlockIdx_out = BlockIdx + BlockShift;
        if (BlockIdx_out < 0 || BlockIdx_out >= num_sectors) { flg = false; return -1; }
        flg = true;
        return qn_offset[BlockIdx_out];
    }
    void OpBlockToGlobalRange(Int BlockIdx, Int BlockShift, Int& start, Int& end, bool& flg) const {
        if (BlockIdx < 0 || BlockIdx >= num_sectors) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "BlockIdx out of bounds");
        Int BlockIdx_out = BlockIdx + BlockShift;
        if (BlockIdx_out < 0 || BlockIdx_out >= num_sectors) { flg = false; return; }
        flg = true;
        start = qn_offset[BlockIdx_out];
        end = qn_offset[BlockIdx_out + 1];
    }
    /* src/QuantumNumbers.cpp:124-142 */
    Int GlobalIdxToBlockIdx(Int GlobIdx) const {
        if (GlobIdx < 0 || GlobIdx >= num_states) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "GlobIdx out of bounds");
        Int BlockIdx = -1;
        while (GlobIdx >= qn_offset[BlockIdx + 1]) ++BlockIdx;
        return BlockIdx;
    }
    /* src/QuantumNumbers.cpp:192-201 */
    Int BlockIdxToGlobalIdx(Int BlockIdx, Int LocIdx) const {
        assert(initialized);
        assert((0 <= BlockIdx) && (BlockIdx < num_sectors));
        return qn_offset[BlockIdx] + LocIdx;
    }
};

/* ------------------------------------------------------------------------------------------------
 *  Sequential AIJ matrix — stands in for PETSc Mat (MPIAIJ on one rank).  Rows keep sorted column
 *  indices and keep explicitly-inserted zeros, like MatSetValues(INSERT_VALUES) + assembly.
 * ---------------------------------------------------------------------------------------------- */
struct CSR {
    Int nrows = 0, ncols = 0;
    std::vector<Int> rowptr, col;
    std::vector<Real> val;
    CSR() {}
    CSR(Int m, Int n) : nrows(m), ncols(n), rowptr(m + 1, 0) {}
    Int nnz() const { return (Int)col.size(); }
    void getrow(Int r, Int& nz, const Int*& idx, const Real*& v) const {
        nz = rowptr[r + 1] - rowptr[r];
        idx = col.data() + rowptr[r];
        v = val.data() + rowptr[r];
    }
    /* Build from per-row maps (col -> value) */
    static CSR FromRows(Int m, Int n, const std::vector<std::map<Int, Real>>& rows) {
        CSR A(m, n);
        for (Int r = 0; r < m; ++r) {
            for (auto& kv : rows[r]) { A.col.push_back(kv.first); A.val.push_back(kv.second); }
            A.rowptr[r + 1] = (Int)A.col.size();
        }
        return A;
    }
    CSR Transpose() const { /* MatHermitianTranspose, real scalars */
        CSR T(ncols, nrows);
        std::vector<Int> cnt(ncols + 1, 0);
        for (Int c : col) cnt[c + 1]++;
        for (Int i = 0; i < ncols; ++i) cnt[i + 1] += cnt[i];
        T.rowptr = cnt;
        T.col.resize(col.size());
        T.val.resize(col.size());
        std::vector<Int> pos(cnt.begin(), cnt.end() - 1);
        for (Int r = 0; r < nrows; ++r)
            for (Int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
                Int p = pos[col[k]]++;
                T.col[p] = r;
                T.val[p] = val[k];
            }
        return T;
    }
    std::vector<Real> ToDense() const {
        std::vector<Real> D((size_t)nrows * ncols, 0.0);
        for (Int r = 0; r < nrows; ++r)
            for (Int k = rowptr[r]; k < rowptr[r + 1]; ++k) D[(size_t)r * ncols + col[k]] += val[k];
        return D;
    }
};

/* ------------------------------------------------------------------------------------------------
 *  Block::SpinBase — include/DMRGBlock.hpp:79-434, src/DMRGBlock.cpp
 * ---------------------------------------------------------------------------------------------- */
struct Block {
    bool init = false, init_Sm = false;
    Int num_sites = 0, num_states = 0;
    int spin_twice = 1; /* 1: spin-1/2 (default), 2: spin-1   (src/DMRGBlock.cpp:54-94) */
    QuantumNumbers Magnetization;
    std::vector<CSR> SzData, SpData, SmData;
    CSR H;

    Int NumSites() const { return num_sites; }
    Int NumStates() const { return num_states; }
    bool Initialized() const { return init; }
    Int loc_dim() const { return spin_twice == 1 ? 2 : 3; }

    /* src/DMRGBlock.cpp:44-170 with num_states_in = PETSC_DEFAULT and num_sites == 1;
       single-site operators: src/DMRGBlock.cpp:1106-1225 */
    void InitializeSingleSite(int spin_twice_in = 1) {
        spin_twice = spin_twice_in;
        num_sites = 1;
        num_states = loc_dim();
        SzData.assign(1, CSR());
        SpData.assign(1, CSR());
        SmData.assign(1, CSR());
        std::vector<std::map<Int, Real>> sz(num_states), sp(num_states), h(num_states);
        if (spin_twice == 1) {
            sz[0][0] = +0.5; sz[1][1] = -0.5;      /* :1131-1136 */
            sp[0][1] = +1.0;                        /* :1193-1195 */
            Magnetization.Initialize({+0.5, -0.5}, {1, 1}); /* include/DMRGBlock.hpp:100-106 */
        } else {
            sz[0][0] = +1.0; sz[2][2] = -1.0;       /* :1151-1156 */
            sp[0][1] = std::sqrt(2.0); sp[1][2] = std::sqrt(2.0); /* :1210-1215 */
            Magnetization.Initialize({+1.0, 0.0, -1.0}, {1, 1, 1});
        }
        SzData[0] = CSR::FromRows(num_states, num_states, sz);
        SpData[0] = CSR::FromRows(num_states, num_states, sp);
        H = CSR::FromRows(num_states, num_states, h); /* zero single-site Hamiltonian, :125-127 */
        init = true;
        CheckSectors();
    }
    /* src/DMRGBlock.cpp:173-196: sizes from sector lists, empty operators of the right size */
    void Initialize(Int num_sites_in, const std::vector<Real>& qn_list_in, const std::vector<Int>& qn_size_in) {
        QuantumNumbers tmp;
        tmp.Initialize(qn_list_in, qn_size_in);
        num_sites = num_sites_in;
        num_states = tmp.NumStates();
        SzData.assign(num_sites, CSR(num_states, num_states));
        SpData.assign(num_sites, CSR(num_states, num_states));
        SmData.assign(num_sites, CSR());
        H = CSR(num_states, num_states);
        Magnetization = tmp;
        init = true;
        init_Sm = false;
    }
    const CSR& Sz(Int i) const { if (i < 0 || i >= num_sites) throw std::runtime_error("Attempted to access non-existent site."); return SzData[i]; }
    const CSR& Sp(Int i) const { if (i < 0 || i >= num_sites) throw std::runtime_error("Attempted to access non-existent site."); return SpData[i]; }
    const CSR& Sm(Int i) const {
        if (i < 0 || i >= num_sites) throw std::runtime_error("Attempted to access non-existent site.");
        if (!init_Sm) throw std::runtime_error("Sm matrices were not initialized. Call CreateSm() first.");
        return SmData[i];
    }
    /* src/DMRGBlock.cpp:623-636 */
    void CreateSm() {
        if (init_Sm) ORACLE_THROW(1, "Sm was previously initialized. Call DestroySm() first.");
        for (Int i = 0; i < num_sites; ++i) SmData[i] = SpData[i].Transpose();
        init_Sm = true;
    }
    void DestroySm() { for (auto& m : SmData) m = CSR(); init_Sm = false; }

    /* src/DMRGBlock.cpp:375-411 */
    void CheckOperatorArray(Op_t OpType) const {
        const std::vector<CSR>* Op;
        switch (OpType) {
            case OpSm: Op = &SmData; break;
            case OpSz: Op = &SzData; break;
            case OpSp: Op = &SpData; break;
            default: ORACLE_THROW(ERR_ARG_WRONG, "Incorrect operator type.");
        }
        for (Int isite = 0; isite < num_sites; ++isite) {
            const CSR& M = (*Op)[isite];
            if (M.rowptr.empty()) ORACLE_THROW(ERR_ARG_CORRUPT, "matrix not yet created.");
            if (M.nrows != M.ncols) ORACLE_THROW(ERR_ARG_WRONG, "matrix not square.");
            if (M.nrows != num_states) ORACLE_THROW(ERR_ARG_WRONG, "matrix dimension does not match the number of states.");
        }
    }
    /* src/DMRGBlock.cpp:413-426 */
    void CheckOperators() const {
        if (!init) ORACLE_THROW(ERR_ARG_CORRUPT, "Block not yet initialized");
        CheckOperatorArray(OpSz);
        CheckOperatorArray(OpSp);
        if (init_Sm) CheckOperatorArray(OpSm);
    }
    /* src/DMRGBlock.cpp:428-447 */
    void CheckSectors() const {
        if (!Magnetization.initialized) ORACLE_THROW(ERR_ARG_WRONGSTATE, "Magnetization not initialized");
        if (num_states != Magnetization.NumStates())
            ORACLE_THROW(ERR_ARG_WRONG, "The number of states in the Magnetization object and the internal value do not match.");
    }
    /* src/DMRGBlock.cpp:517-600: only the first and last column of every row are range-checked */
    void MatCheckOperatorBlocks(Op_t OpType, const CSR& matin) const {
        CheckSectors();
        if (Magnetization.NumStates() != matin.nrows) ORACLE_THROW(1, "Incorrect number of rows.");
        for (Int row = 0; row < matin.nrows; ++row) {
            Int blk = Magnetization.GlobalIdxToBlockIdx(row);
            Int cs = 0, ce = 0;
            bool flg;
            Magnetization.OpBlockToGlobalRange(blk, (Int)OpType, cs, ce, flg);
            Int nz; const Int* c; const Real* v;
            matin.getrow(row, nz, c, v);
            /* (the reference's "should have no entries" test requires nzA!=0 && nzB!=0, which a
               one-rank matrix never satisfies — :575; the index check below then fires instead) */
            if (nz) {
                if (!flg) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "Row should have no entries.");
                if (c[0] < cs || c[0] >= ce) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "column index out of the sector block");
                if (c[nz - 1] < cs || c[nz - 1] >= ce) ORACLE_THROW(ERR_ARG_OUTOFRANGE, "column index out of the sector block");
            }
        }
    }
    /* src/DMRGBlock.cpp:603-620 */
    void CheckOperatorBlocks() const {
        if (!init) ORACLE_THROW(ERR_ARG_CORRUPT, "Block not yet initialized");
        CheckOperators();
        for (Int i = 0; i < num_sites; ++i) MatCheckOperatorBlocks(OpSz, SzData[i]);
        for (Int i = 0; i < num_sites; ++i) MatCheckOperatorBlocks(OpSp, SpData[i]);
    }
};

/* ------------------------------------------------------------------------------------------------
 *  Hamiltonians::J1J2XXZModel_SquareLattice — include/Hamiltonians.hpp:77-288, src/Hamiltonians.cpp
 * ---------------------------------------------------------------------------------------------- */
struct Term { Real a; Op_t Iop; Int Isite; Op_t Jop; Int Jsite; }; /* include/Hamiltonians.hpp:17-24 */
enum BC_t { OpenBC = 0, PeriodicBC = 1 };

struct Hamiltonian {
    Real J1 = 1.0, Jz1 = 0.0, J2 = 1.0, Jz2 = 0.0; /* defaults: include/Hamiltonians.hpp:240-252 */
    Int Lx = 4, Ly = 4;
    BC_t BCx = OpenBC, BCy = PeriodicBC;
    bool heisenberg = false;
    /* include/Hamiltonians.hpp:102-108 */
    void SetHeisenberg(Real Jz) { heisenberg = true; Jz1 = Jz; J1 = 0.5; J2 = 0.0; Jz2 = 0.0; }
    Int NumSites() const { return Lx * Ly; }
    Int NumEnvSites() const { return Ly; }
    /* src/Hamiltonians.cpp:4 — s-shaped snake */
    Int To1D(Int ix, Int jy) const { return (ix * Ly + jy) * (1 - (ix % 2)) + ((ix + 1) * Ly - (jy + 1)) * (ix % 2); }
    /* src/Hamiltonians.cpp:26-48 */
    std::vector<Int> GetNearestNeighbors(Int ix, Int jy, Int nsites_in) const {
        std::vector<Int> nn;
        if (((0 <= jy) && (jy < (Ly - 1))) || ((jy == (Ly - 1)) && (BCy == PeriodicBC))) {
            const Int jy_above = (jy + 1) % Ly;
            const Int nn1d = To1D(ix, jy_above);
            if (nn1d < nsites_in && jy_above != jy) nn.push_back(nn1d);
        }
        if (((0 <= ix) && (ix < (Lx - 1))) || ((ix == (Lx - 1)) && (BCx == PeriodicBC))) {
            const Int ix_right = (ix + 1) % Lx;
            const Int nn1d = To1D(ix_right, jy);
            if (nn1d < nsites_in && ix_right != ix) nn.push_back(nn1d);
        }
        return nn;
    }
    /* src/Hamiltonians.cpp:50-71 */
    std::vector<Int> GetNextNearestNeighbors(Int ix, Int jy, Int nsites_in) const {
        std::vector<Int> nnn;
        if ((((1 <= ix) && (ix < Lx)) || ((ix == 0) && (BCx == PeriodicBC))) &&
            (((0 <= jy) && (jy < Ly - 1)) || ((jy == (Ly - 1)) && (BCy == PeriodicBC)))) {
            const Int n1 = To1D((ix + Lx - 1) % Lx, (jy + 1) % Ly);
            if (n1 < nsites_in) nnn.push_back(n1);
        }
        if ((((0 <= ix) && (ix < Lx - 1)) || ((ix == (Lx - 1)) && (BCx == PeriodicBC))) &&
            (((0 <= jy) && (jy < Ly - 1)) || ((jy == (Ly - 1)) && (BCy == PeriodicBC)))) {
            const Int n1 = To1D((ix + 1) % Lx, (jy + 1) % Ly);
            if (n1 < nsites_in) nnn.push_back(n1);
        }
        return nnn;
    }
    /* src/Hamiltonians.cpp:73-122 (nsites_in < 0 == PETSC_DEFAULT) */
    std::vector<Term> H(Int nsites_in) const {
        Int ns = (nsites_in < 0) ? Lx * Ly : nsites_in;
        std::vector<Term> Terms;
        for (Int is = 0; is < ns; ++is) {
            const Int ix = is / Ly;
            const Int jy = (is % Ly) * (1 - 2 * (ix % 2)) + (Ly - 1) * (ix % 2);
            if (J1 != 0.0 || Jz1 != 0.0) {
                for (Int in : GetNearestNeighbors(ix, jy, ns)) {
                    Int ia = (in < is) ? in : is;
                    Int ib = (in > is) ? in : is;
                    if (J1 != 0.0) Terms.push_back({J1, OpSp, ia, OpSm, ib});
                    if (J1 != 0.0) Terms.push_back({J1, OpSm, ia, OpSp, ib});
                    if (Jz1 != 0.0) Terms.push_back({Jz1, OpSz, ia, OpSz, ib});
                }
            }
            /* NNN terms only when BOTH J2 and Jz2 are non-zero (:101) */
            if ((J2 != 0.0 && Jz2 != 0.0) && Lx > 1 && Ly > 1) {
                for (Int in : GetNextNearestNeighbors(ix, jy, ns)) {
                    Int il = (in < is) ? in : is;
                    Int ir = (in > is) ? in : is;
                    if (J2 != 0.0) Terms.push_back({J2, OpSp, il, OpSm, ir});
                    if (J2 != 0.0) Terms.push_back({J2, OpSm, il, OpSp, ir});
                    if (Jz2 != 0.0) Terms.push_back({Jz2, OpSz, il, OpSz, ir});
                }
            }
        }
        return Terms;
    }
};

/* ------------------------------------------------------------------------------------------------
 *  KronBlocks_t — include/DMRGKron.hpp:117-480
 * ---------------------------------------------------------------------------------------------- */
typedef std::tuple<Real, Int, Int, Int> KronBlock_t; /* include/DMRGKron.hpp:22 */

struct KronSumTerm { Real a; Op_t OpTypeA; const CSR* A; Op_t OpTypeB; const CSR* B; }; /* :28-34 */

/* include/DMRGKron.hpp:85-89 — 72 bytes per (row, term) in the reference */
struct KronSumTermRow {
    Int nz_L, nz_R, bks_L, col_NStatesR, fws_O;
    const Int *idx_L, *idx_R;
    const Real *v_L, *v_R;
};

struct KronBlocks_t;

/* The shell context — include/DMRGKron.hpp:92-112 (one rank: rstart=0, lrows=Nrows unless split) */
struct KronSumShell {
    Int rstart = 0, rend = 0, lrows = 0, Nrows = 0;
    std::vector<KronSumTerm> Terms;
    std::vector<KronSumTermRow> kstr;
    Int Nterms = 0;
    std::vector<Int> Rows_L, Rows_R;
    std::vector<Real> term_a;
    Real one = 1.0;
    /* operands kept alive for the lifetime of the shell (the reference keeps SeqAIJ submatrices) */
    std::vector<CSR> owned;
    /* src/DMRGKron.cpp:1827-1869, rows [r0,r1) of the local range (for the threaded CPU baseline) */
    void MatMultRows(const Real* xvals, Real* yvals, Int r0, Int r1) const {
        for (Int ir = r0; ir < r1; ++ir) {
            Real yval = 0.0;
            Int irt = ir * Nterms - 1;
            for (Int it = 0; it < Nterms; ++it) {
                ++irt;
                const KronSumTermRow& k = kstr[irt];
                for (Int l = 0; l < k.nz_L; ++l) {
                    Int idx = (k.idx_L[l] - k.bks_L) * k.col_NStatesR + k.fws_O;
                    Real temp = term_a[it] * k.v_L[l];
                    for (Int r = 0; r < k.nz_R; ++r) yval += temp * k.v_R[r] * xvals[idx + k.idx_R[r]];
                }
            }
            yvals[ir] = yval;
        }
    }
    void MatMult(const Real* x, Real* y) const { MatMultRows(x, y, 0, lrows); }
    /* Σ_rows Σ_terms nz_L*nz_R — the reference's multiply-add count per apply */
    double UnfactoredFMAs() const {
        double s = 0;
        for (auto& k : kstr) s += (double)k.nz_L * (double)k.nz_R;
        return s;
    }
};

struct KronBlocks_t {
    const Block& LeftBlock;
    const Block& RightBlock;
    std::vector<KronBlock_t> KronBlocks;
    std::vector<Real> kb_list;
    std::vector<Int> kb_size, kb_offset;
    std::map<std::tuple<Int, Int>, Int> kb_map;
    Int num_blocks = 0, num_states = 0;
    Real ks_tol = 1.0e-16; /* include/DMRGKron.hpp:396 */

    /* include/DMRGKron.hpp:124-213 */
    KronBlocks_t(const Block& L, const Block& R, const std::vector<Real>& QNSectors) : LeftBlock(L), RightBlock(R) {
        if (!L.Initialized()) throw std::runtime_error("Left input block not initialized.");
        if (!R.Initialized()) throw std::runtime_error("Right input block not initialized.");
        const auto& LL = L.Magnetization.List();
        const auto& RL = R.Magnetization.List();
        const auto& LS = L.Magnetization.Sizes();
        const auto& RS = R.Magnetization.Sizes();
        if (QNSectors.size() == 0) {
            for (size_t IL = 0; IL < LL.size(); ++IL)
                for (size_t IR = 0; IR < RL.size(); ++IR)
                    KronBlocks.push_back(std::make_tuple(LL[IL] + RL[IR], (Int)IL, (Int)IR, LS[IL] * RS[IR]));
            std::stable_sort(KronBlocks.begin(), KronBlocks.end(),
                             [](const KronBlock_t& a, const KronBlock_t& b) { return std::get<0>(a) > std::get<0>(b); });
        } else if (QNSectors.size() == 1) {
            for (size_t IL = 0; IL < LL.size(); ++IL)
                for (size_t IR = 0; IR < RL.size(); ++IR) {
                    Real QN = LL[IL] + RL[IR];
                    if (QN == QNSectors[0]) KronBlocks.push_back(std::make_tuple(QN, (Int)IL, (Int)IR, LS[IL] * RS[IR]));
                }
        } else {
            std::set<Real> S(QNSectors.begin(), QNSectors.end());
            for (size_t IL = 0; IL < LL.size(); ++IL)
                for (size_t IR = 0; IR < RL.size(); ++IR) {
                    Real QN = LL[IL] + RL[IR];
                    if (S.find(QN) != S.end()) KronBlocks.push_back(std::make_tuple(QN, (Int)IL, (Int)IR, LS[IL] * RS[IR]));
                }
        }
        num_blocks = (Int)KronBlocks.size();
        for (auto& kb : KronBlocks) kb_list.push_back(std::get<0>(kb));
        for (auto& kb : KronBlocks) kb_size.push_back(std::get<3>(kb));
        Int idx = 0;
        for (auto& kb : KronBlocks) kb_map[std::make_tuple(std::get<1>(kb), std::get<2>(kb))] = idx++;
        Int sum = 0;
        for (auto& kb : KronBlocks) { kb_offset.push_back(sum); sum += std::get<3>(kb); }
        kb_offset.push_back(sum);
        num_states = sum;
    }
    Int size() const { return (Int)KronBlocks.size(); }
    const std::vector<KronBlock_t>& data() const { return KronBlocks; }
    Real QN(Int i) const { return std::get<0>(KronBlocks[i]); }
    Int LeftIdx(Int i) const { return std::get<1>(KronBlocks[i]); }
    Int RightIdx(Int i) const { return std::get<2>(KronBlocks[i]); }
    Int Sizes(Int i) const { return std::get<3>(KronBlocks[i]); }
    Int Offsets(Int i) const { assert(i >= 0 && i < num_blocks + 1); return kb_offset[i]; }
    /* include/DMRGKron.hpp:283-294 */
    Int Map(Int l, Int r) const { auto f = kb_map.find(std::make_tuple(l, r)); return f != kb_map.end() ? f->second : -1; }
    /* include/DMRGKron.hpp:272-276 */
    Int Offsets(Int l, Int r) const { Int i = Map(l, r); return i >= 0 ? kb_offset[i] : -1; }
    Int NumStates() const { return num_states; }

    void KronSumConstructExplicit(const Block& L, const Block& R, const std::vector<Term>& TermsLR, CSR& MatOut) const;
    void KronSumConstruct(Block& L, Block& R, const std::vector<Term>& Terms, CSR* MatOutExplicit, KronSumShell* shell,
                          Int rstart = -1, Int rend = -1) const;
    void KronSumSetUpShellTerms(KronSumShell& sh) const;
    void BuildTerms(const Block& L, const Block& R, const std::vector<Term>& TermsLR, std::vector<KronSumTerm>& out,
                    std::vector<CSR>& owned) const;
};

/* include/DMRGKron.hpp:501-656 */
struct KronBlocksIterator {
    const KronBlocks_t& KB;
    Int istart_, iend_, idx_, blockidx_ = -1;
    std::vector<Int> kb_size, kb_offset;
    Int num_states = 0;
    bool updated_block = true;
    KronBlocksIterator(const KronBlocks_t& KB_, Int s, Int e) : KB(KB_), istart_(s), iend_(e), idx_(s) {
        if (istart_ == iend_) return;
        Int sum = 0;
        for (auto& kb : KB.data()) { kb_size.push_back(std::get<3>(kb)); kb_offset.push_back(sum); sum += std::get<3>(kb); }
        kb_offset.push_back(sum);
        num_states = sum;
        assert(istart_ < sum);
        while (idx_ >= kb_offset[blockidx_ + 1]) ++blockidx_;
    }
    bool Loop() const { return idx_ < iend_; }
    Int Steps() const { return idx_ - istart_; }
    Int BlockIdx() const { return blockidx_; }
    Int LocIdx() const { return idx_ - kb_offset[blockidx_]; }
    /* :555-566 — note the bound is num_states, not the block count */
    Int BlockStartIdx(Int BlockShift) const {
        Int o = blockidx_ + BlockShift;
        if (o < 0 || o >= num_states) return -1;
        return kb_offset[o];
    }
    void operator++() {
        ++idx_;
        if (idx_ >= kb_offset[blockidx_ + 1]) { ++blockidx_; updated_block = true; } else updated_block = false;
    }
    Int BlockIdxLeft() const { return std::get<1>(KB.data()[blockidx_]); }
    Int BlockIdxRight() const { return std::get<2>(KB.data()[blockidx_]); }
    Int NumStatesRight() const { return KB.RightBlock.Magnetization.Sizes()[BlockIdxRight()]; }
    Int LocIdxLeft() const { return LocIdx() / NumStatesRight(); }
    Int LocIdxRight() const { return LocIdx() % NumStatesRight(); }
    Int GlobalIdxLeft() const { return KB.LeftBlock.Magnetization.BlockIdxToGlobalIdx(BlockIdxLeft(), LocIdxLeft()); }
    Int GlobalIdxRight() const { return KB.RightBlock.Magnetization.BlockIdxToGlobalIdx(BlockIdxRight(), LocIdxRight()); }
    bool UpdatedBlock() const { return updated_block; }
};

/* src/DMRGKron.cpp:891-989 (KronSumGetSubmatrices): term list = {H_L⊗1, 1⊗H_R, LR terms in input order}.
   On one rank the "submatrices" are the operators themselves. */
inline void KronBlocks_t::BuildTerms(const Block& L, const Block& R, const std::vector<Term>& TermsLR,
                                     std::vector<KronSumTerm>& out, std::vector<CSR>& owned) const {
    (void)owned;
    out.clear();
    out.push_back({1.0, OpSz, &L.H, OpEye, nullptr});
    out.push_back({1.0, OpEye, nullptr, OpSz, &R.H});
    auto get = [](const Block& B, Op_t op, Int isite) -> const CSR* {
        return op == OpSp ? &B.Sp(isite) : (op == OpSm ? &B.Sm(isite) : (op == OpSz ? &B.Sz(isite) : nullptr));
    };
    for (const Term& t : TermsLR) out.push_back({t.a, t.Iop, get(L, t.Iop, t.Isite), t.Jop, get(R, t.Jop, t.Jsite)});
}

/* src/DMRGKron.cpp:1706-1824 */
inline void KronBlocks_t::KronSumSetUpShellTerms(KronSumShell& sh) const {
    sh.Nterms = (Int)sh.Terms.size();
    sh.Rows_L.assign(sh.lrows, 0);
    sh.Rows_R.assign(sh.lrows, 0);
    sh.kstr.assign((size_t)sh.Nterms * sh.lrows, KronSumTermRow());
    sh.term_a.resize(sh.Nterms);
    for (Int it = 0; it < sh.Nterms; ++it) sh.term_a[it] = sh.Terms[it].a;
    std::map<Op_t, Int> fws_LOP, Row_NumStates_ROP;
    Int irt = 0;
    if (sh.lrows > 0) {
        KronBlocksIterator KIter(*this, sh.rstart, sh.rend);
        for (; KIter.Loop(); ++KIter) {
            const Int lrow = KIter.Steps();
            const Int Row_BlockIdx_L = KIter.BlockIdxLeft();
            const Int Row_BlockIdx_R = KIter.BlockIdxRight();
            const Int Row_L = sh.Rows_L[lrow] = KIter.GlobalIdxLeft();
            const Int Row_R = sh.Rows_R[lrow] = KIter.GlobalIdxRight();
            bool flg[2];
            if (KIter.UpdatedBlock()) {
                fws_LOP = {{OpEye, KIter.BlockStartIdx(OpSz)},
                           {OpSz, KIter.BlockStartIdx(OpSz)},
                           {OpSp, Offsets(Row_BlockIdx_L + 1, Row_BlockIdx_R - 1)},
                           {OpSm, Offsets(Row_BlockIdx_L - 1, Row_BlockIdx_R + 1)}};
                Row_NumStates_ROP = {{OpEye, KIter.NumStatesRight()},
                                     {OpSz, KIter.NumStatesRight()},
                                     {OpSp, RightBlock.Magnetization.Sizes(Row_BlockIdx_R + 1)},
                                     {OpSm, RightBlock.Magnetization.Sizes(Row_BlockIdx_R - 1)}};
            }
            for (const KronSumTerm& term : sh.Terms) {
                KronSumTermRow& k = sh.kstr[irt];
                Int bks_R = 0;
                if (term.OpTypeA != OpEye) {
                    term.A->getrow(Row_L, k.nz_L, k.idx_L, k.v_L);
                    k.bks_L = LeftBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_L, term.OpTypeA, flg[SideLeft]);
                } else {
                    k.nz_L = 1; k.idx_L = &sh.Rows_L[lrow]; k.v_L = &sh.one;
                    k.bks_L = LeftBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_L, OpSz, flg[SideLeft]);
                }
                if (term.OpTypeB != OpEye) {
                    term.B->getrow(Row_R, k.nz_R, k.idx_R, k.v_R);
                    bks_R = RightBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_R, term.OpTypeB, flg[SideRight]);
                } else {
                    k.nz_R = 1; k.idx_R = &sh.Rows_R[lrow]; k.v_R = &sh.one;
                    bks_R = RightBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_R, OpSz, flg[SideRight]);
                }
                if ((!flg[SideLeft]) || (!flg[SideRight]) || (k.nz_L * k.nz_R == 0)) {
                    k.nz_L = 0; k.nz_R = 0;
                } else {
                    k.fws_O = fws_LOP.at(term.OpTypeA) - bks_R;
                    k.col_NStatesR = Row_NumStates_ROP.at(term.OpTypeB);
                    if (k.col_NStatesR == -1) ORACLE_THROW(1, "Accessed incorrect value.");
                }
                ++irt;
            }
        }
    }
}

/* src/DMRGKron.cpp:844-881 + 1340-1477 (KronSumFillMatrix): explicit Σ a A⊗B restricted to the
   KronBlocks basis, |v| < ks_tol filtered.  The reference inserts a dense row span (with zeros);
   PETSc's AIJ assembly keeps whatever was preallocated — only the non-zero values matter for
   everything downstream, so zeros are dropped here (documented deviation in structure only). */
inline void KronBlocks_t::KronSumConstructExplicit(const Block& L, const Block& R, const std::vector<Term>& TermsLR,
                                                   CSR& MatOut) const {
    std::vector<KronSumTerm> Terms;
    std::vector<CSR> owned;
    BuildTerms(L, R, TermsLR, Terms, owned);
    const Int N = num_states;
    MatOut = CSR(N, N);
    std::vector<Real> val_arr(N, 0.0);
    std::vector<Int> touched;
    std::map<Op_t, Int> fws_LOP, Row_NumStates_ROP;
    KronBlocksIterator KIter(*this, 0, N);
    for (; KIter.Loop(); ++KIter) {
        const Int Irow = KIter.Steps();
        const Int Row_BlockIdx_L = KIter.BlockIdxLeft();
        const Int Row_BlockIdx_R = KIter.BlockIdxRight();
        const Int Row_L = KIter.GlobalIdxLeft();
        const Int Row_R = KIter.GlobalIdxRight();
        bool flg[2];
        Int nz_L, nz_R, bks_L, bks_R, col_NStatesR, fws_O;
        const Int *idx_L, *idx_R;
        const Real *v_L, *v_R;
        const Real one = 1.0;
        if (KIter.UpdatedBlock()) {
            fws_LOP = {{OpEye, KIter.BlockStartIdx(OpSz)},
                       {OpSz, KIter.BlockStartIdx(OpSz)},
                       {OpSp, Offsets(Row_BlockIdx_L + 1, Row_BlockIdx_R - 1)},
                       {OpSm, Offsets(Row_BlockIdx_L - 1, Row_BlockIdx_R + 1)}};
            Row_NumStates_ROP = {{OpEye, KIter.NumStatesRight()},
                                 {OpSz, KIter.NumStatesRight()},
                                 {OpSp, RightBlock.Magnetization.Sizes(Row_BlockIdx_R + 1)},
                                 {OpSm, RightBlock.Magnetization.Sizes(Row_BlockIdx_R - 1)}};
        }
        touched.clear();
        for (const KronSumTerm& term : Terms) {
            if (term.OpTypeA != OpEye) {
                term.A->getrow(Row_L, nz_L, idx_L, v_L);
                bks_L = LeftBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_L, term.OpTypeA, flg[SideLeft]);
            } else {
                nz_L = 1; idx_L = &Row_L; v_L = &one;
                bks_L = LeftBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_L, OpSz, flg[SideLeft]);
            }
            if (term.OpTypeB != OpEye) {
                term.B->getrow(Row_R, nz_R, idx_R, v_R);
                bks_R = RightBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_R, term.OpTypeB, flg[SideRight]);
            } else {
                nz_R = 1; idx_R = &Row_R; v_R = &one;
                bks_R = RightBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_R, OpSz, flg[SideRight]);
            }
            if (!(flg[SideLeft] && flg[SideRight])) continue;
            if (nz_L * nz_R == 0) continue;
            fws_O = fws_LOP.at(term.OpTypeA);
            col_NStatesR = Row_NumStates_ROP.at(term.OpTypeB);
            if (col_NStatesR == -1) ORACLE_THROW(1, "Accessed incorrect value.");
            for (Int l = 0; l < nz_L; ++l)
                for (Int r = 0; r < nz_R; ++r) {
                    Int c = (idx_L[l] - bks_L) * col_NStatesR + (idx_R[r] - bks_R) + fws_O;
                    val_arr[c] += term.a * v_L[l] * v_R[r];
                    touched.push_back(c);
                }
        }
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        for (Int c : touched) {
            if (!(std::fabs(val_arr[c]) < ks_tol)) { MatOut.col.push_back(c); MatOut.val.push_back(val_arr[c]); }
            val_arr[c] = 0.0;
        }
        MatOut.rowptr[Irow + 1] = (Int)MatOut.col.size();
    }
}

/* src/DMRGKron.cpp:759-841.  Exactly one of MatOutExplicit / shell is non-null (do_shell switch). */
inline void KronBlocks_t::KronSumConstruct(Block& L, Block& R, const std::vector<Term>& Terms, CSR* MatOutExplicit,
                                           KronSumShell* shell, Int rstart, Int rend) const {
    const Int nsites_left = L.NumSites(), nsites_right = R.NumSites(), nsites_out = nsites_left + nsites_right;
    Int Max_Isite = 0;
    for (const Term& t : Terms) { Max_Isite = std::max(Max_Isite, t.Isite); Max_Isite = std::max(Max_Isite, t.Jsite); }
    if (Max_Isite >= nsites_out) ORACLE_THROW(1, "Maximum site index from Terms has to be less than the total number of sites.");
    L.CheckOperators(); L.CheckSectors(); L.CheckOperatorBlocks();
    R.CheckOperators(); R.CheckSectors(); R.CheckOperatorBlocks();
    std::vector<Term> TermsLR;
    for (const Term& t : Terms) {
        if ((0 <= t.Isite && t.Isite < nsites_left) && (nsites_left <= t.Jsite && t.Jsite < nsites_out)) {
            if (t.a == 0.0) continue;
            TermsLR.push_back(t);
        } else if ((0 <= t.Isite && t.Isite < nsites_left) && (0 <= t.Jsite && t.Jsite < nsites_left)) {
        } else if ((nsites_left <= t.Isite && t.Isite < nsites_out) && (nsites_left <= t.Jsite && t.Jsite < nsites_out)) {
        } else ORACLE_THROW(1, "Invalid term.");
    }
    /* reflection: new sites always at the interface (:803-807) */
    for (Term& t : TermsLR) t.Jsite = nsites_out - 1 - t.Jsite;
    bool CreateSmL = false, CreateSmR = false;
    for (const Term& t : TermsLR) if (t.Iop == OpSm) { CreateSmL = true; break; }
    for (const Term& t : TermsLR) if (t.Jop == OpSm) { CreateSmR = true; break; }
    /* the same object may be passed as L and R (SysBlockEnl == EnvBlockEnl) */
    if (CreateSmL && !L.init_Sm) L.CreateSm();
    if (CreateSmR && !R.init_Sm) R.CreateSm();
    if (shell) {
        /* src/DMRGKron.cpp:1871-1917; [rstart,rend) is this "rank"'s row range (default: one rank owns all rows) */
        shell->Nrows = num_states;
        shell->rstart = rstart < 0 ? 0 : rstart;
        shell->rend = rend < 0 ? num_states : rend;
        shell->lrows = shell->rend - shell->rstart;
        BuildTerms(L, R, TermsLR, shell->Terms, shell->owned);
        KronSumSetUpShellTerms(*shell);
        /* the shell aliases operator rows, so Sm must outlive it: the caller destroys Sm after use */
    } else {
        KronSumConstructExplicit(L, R, TermsLR, *MatOutExplicit);
        if (CreateSmL) L.DestroySm();
        if (CreateSmR && R.init_Sm) R.DestroySm();
    }
}

/* ------------------------------------------------------------------------------------------------
 *  MatKronEyeConstruct / KronEye_Explicit — src/DMRGKron.cpp:52-456, 459-615
 * ---------------------------------------------------------------------------------------------- */
inline void KronEye_Explicit(Block& LeftBlock, Block& RightBlock, const std::vector<Term>& Terms, Block& BlockOut) {
    if (!LeftBlock.Initialized()) ORACLE_THROW(1, "Left input block not initialized.");
    if (!RightBlock.Initialized()) ORACLE_THROW(1, "Right input block not initialized.");
    LeftBlock.CheckOperators(); LeftBlock.CheckSectors(); LeftBlock.CheckOperatorBlocks();
    RightBlock.CheckOperators(); RightBlock.CheckSectors(); RightBlock.CheckOperatorBlocks();
    KronBlocks_t KronBlocks(LeftBlock, RightBlock, {});
    const Int nsites_left = LeftBlock.NumSites(), nsites_right = RightBlock.NumSites();
    const Int nsites_out = nsites_left + nsites_right;
    const Int nstates_out = LeftBlock.NumStates() * RightBlock.NumStates();
    if (KronBlocks.NumStates() != nstates_out) ORACLE_THROW(1, "Mismatch in number of states.");
    /* merge equal-QN KronBlocks into sectors (:560-574) */
    std::vector<Real> QN_List;
    std::vector<Int> QN_Size;
    Real QN_last = 0;
    for (auto& tup : KronBlocks.data()) {
        const Real qn = std::get<0>(tup);
        const Int size = std::get<3>(tup);
        if (qn < QN_last || QN_List.size() == 0) { QN_List.push_back(qn); QN_Size.push_back(size); }
        else QN_Size.back() += size;
        QN_last = qn;
    }
    BlockOut.Initialize(nsites_out, QN_List, QN_Size);
    BlockOut.spin_twice = LeftBlock.spin_twice;

    /* MatKronEyeConstruct, fill pass (:323-437), one rank: rstart = 0, lrows = nstates_out */
    const Int TotSites = nsites_out;
    const Int SiteShifts_LR[2] = {0, nsites_left};
    const Int NumSites_LR[2] = {nsites_left, nsites_right};
    std::vector<std::vector<std::map<Int, Real>>> rowsZ(TotSites, std::vector<std::map<Int, Real>>(nstates_out));
    std::vector<std::vector<std::map<Int, Real>>> rowsP(TotSites, std::vector<std::map<Int, Real>>(nstates_out));
    const Real one = 1.0;
    KronBlocksIterator KIter(KronBlocks, 0, nstates_out);
    Int fws_O_Sp_LR[2] = {-1, -1}, col_NStatesR_LR[2] = {-1, -1};
    for (; KIter.Loop(); ++KIter) {
        const Int Irow = KIter.Steps();
        const Int Row_BlockIdx_L = KIter.BlockIdxLeft();
        const Int Row_BlockIdx_R = KIter.BlockIdxRight();
        const Int Row_NumStates_R = KIter.NumStatesRight();
        const Int Row_LocIdx_L = KIter.LocIdxLeft();
        const Int Row_LocIdx_R = KIter.LocIdxRight();
        const Int LocRow_L = KIter.GlobalIdxLeft();
        const Int LocRow_R = KIter.GlobalIdxRight();
        bool flg[2];
        Int nz_L, nz_R, col_NStatesR;
        const Int *idx_L, *idx_R;
        const Real *v_L, *v_R, *v_O;
        const Int fws_O_Sz = KIter.BlockStartIdx(OpSz);
        if (KIter.UpdatedBlock()) {
            fws_O_Sp_LR[0] = KronBlocks.Offsets(Row_BlockIdx_L + 1, Row_BlockIdx_R);
            fws_O_Sp_LR[1] = KronBlocks.Offsets(Row_BlockIdx_L, Row_BlockIdx_R + 1);
            col_NStatesR_LR[0] = RightBlock.Magnetization.Sizes(Row_BlockIdx_R);
            col_NStatesR_LR[1] = RightBlock.Magnetization.Sizes(Row_BlockIdx_R + 1);
        }
        for (Op_t OpType : {OpSz, OpSp}) {
            const Int shift_L[2] = {LeftBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_L, OpType, flg[SideLeft]), 0};
            const Int shift_R[2] = {0, RightBlock.Magnetization.OpBlockToGlobalRangeStart(Row_BlockIdx_R, OpType, flg[SideRight])};
            for (int SideType : {SideLeft, SideRight}) {
                Int fws_O;
                if (OpType == OpSz) { col_NStatesR = Row_NumStates_R; fws_O = fws_O_Sz; if (fws_O == -1) continue; }
                else { col_NStatesR = col_NStatesR_LR[SideType]; fws_O = fws_O_Sp_LR[SideType]; if (fws_O == -1) continue; }
                const Int ishift = SiteShifts_LR[SideType];
                for (Int isite = 0; isite < NumSites_LR[SideType]; ++isite) {
                    if (!flg[SideType]) continue;
                    const Int bks_L = shift_L[SideType];
                    const Int bks_R = shift_R[SideType];
                    if (SideType) { /* right */
                        const CSR& mat = (OpType == OpSz) ? RightBlock.Sz(isite) : RightBlock.Sp(isite);
                        nz_L = 1; idx_L = &Row_LocIdx_L; v_L = &one;
                        mat.getrow(LocRow_R, nz_R, idx_R, v_R);
                        v_O = v_R;
                    } else {
                        const CSR& mat = (OpType == OpSz) ? LeftBlock.Sz(isite) : LeftBlock.Sp(isite);
                        mat.getrow(LocRow_L, nz_L, idx_L, v_L);
                        nz_R = 1; idx_R = &Row_LocIdx_R; v_R = &one;
                        v_O = v_L;
                    }
                    (void)v_L; (void)v_R;
                    auto& row = (OpType == OpSz ? rowsZ : rowsP)[isite + ishift][Irow];
                    for (Int l = 0; l < nz_L; ++l)
                        for (Int r = 0; r < nz_R; ++r)
                            row[(idx_L[l] - bks_L) * col_NStatesR + (idx_R[r] - bks_R) + fws_O] = v_O[l * nz_R + r];
                }
            }
        }
    }
    for (Int i = 0; i < TotSites; ++i) {
        BlockOut.SzData[i] = CSR::FromRows(nstates_out, nstates_out, rowsZ[i]);
        BlockOut.SpData[i] = CSR::FromRows(nstates_out, nstates_out, rowsP[i]);
    }
    /* :604-612 — enlarged-block Hamiltonian through the explicit KronSum (do_shell = FALSE default) */
    KronBlocks.KronSumConstruct(LeftBlock, RightBlock, Terms, &BlockOut.H, nullptr);
}

/* ------------------------------------------------------------------------------------------------
 *  Dense symmetric eigensolver standing in for EPSLAPACK (include/DMRGBlockContainer.hpp:1976-1982):
 *  Householder tridiagonalisation + implicit QL (EISPACK tred2/tql2 algorithm), all eigenpairs,
 *  returned in DESCENDING order (EPS_LARGEST_REAL).  V is n×n row-major, eigenvector k = column k.
 * ---------------------------------------------------------------------------------------------- */
inline void SymEigDescending(Int n, std::vector<Real> A /* row-major, copied */, std::vector<Real>& w, std::vector<Real>& Vout) {
    std::vector<Real>& V = A;
    std::vector<Real> d(n), e(n);
    auto at = [&](Int i, Int j) -> Real& { return V[(size_t)i * n + j]; };
    for (Int j = 0; j < n; j++) d[j] = at(n - 1, j);
    for (Int i = n - 1; i > 0; i--) {
        Real scale = 0.0, h = 0.0;
        for (Int k = 0; k < i; k++) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (Int j = 0; j < i; j++) { d[j] = at(i - 1, j); at(i, j) = 0.0; at(j, i) = 0.0; }
        } else {
            for (Int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
            Real f = d[i - 1];
            Real g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (Int j = 0; j < i; j++) e[j] = 0.0;
            for (Int j = 0; j < i; j++) {
                f = d[j];
                at(j, i) = f;
                g = e[j] + at(j, j) * f;
                for (Int k = j + 1; k <= i - 1; k++) { g += at(k, j) * d[k]; e[k] += at(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (Int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
            Real hh = f / (h + h);
            for (Int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (Int j = 0; j < i; j++) {
                f = d[j]; g = e[j];
                for (Int k = j; k <= i - 1; k++) at(k, j) -= (f * e[k] + g * d[k]);
                d[j] = at(i - 1, j);
                at(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (Int i = 0; i < n - 1; i++) {
        at(n - 1, i) = at(i, i);
        at(i, i) = 1.0;
        Real h = d[i + 1];
        if (h != 0.0) {
            for (Int k = 0; k <= i; k++) d[k] = at(k, i + 1) / h;
            for (Int j = 0; j <= i; j++) {
                Real g = 0.0;
                for (Int k = 0; k <= i; k++) g += at(k, i + 1) * at(k, j);
                for (Int k = 0; k <= i; k++) at(k, j) -= g * d[k];
            }
        }
        for (Int k = 0; k <= i; k++) at(k, i + 1) = 0.0;
    }
    for (Int j = 0; j < n; j++) { d[j] = at(n - 1, j); at(n - 1, j) = 0.0; }
    if (n > 0) at(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
    /* tql2 */
    for (Int i = 1; i < n; i++) e[i - 1] = e[i];
    if (n > 0) e[n - 1] = 0.0;
    Real f = 0.0, tst1 = 0.0;
    const Real eps = std::pow(2.0, -52.0);
    for (Int l = 0; l < n; l++) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        Int m = l;
        while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            int iter = 0;
            do {
                iter++;
                Real g = d[l];
                Real p = (d[l + 1] - g) / (2.0 * e[l]);
                Real r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                Real dl1 = d[l + 1];
                Real h = g - d[l];
                for (Int i = l + 2; i < n; i++) d[i] -= h;
                f += h;
                p = d[m];
                Real c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
                for (Int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (Int k = 0; k < n; k++) {
                        h = at(k, i + 1);
                        at(k, i + 1) = s * at(k, i) + c * h;
                        at(k, i) = c * at(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
                if (iter > 200) throw std::runtime_error("SymEig: no convergence");
            } while (std::fabs(e[l]) > eps * tst1);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    /* sort descending, stable on index */
    std::vector<Int> ord(n);
    for (Int i = 0; i < n; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](Int a, Int b) { return d[a] > d[b]; });
    w.resize(n);
    Vout.assign((size_t)n * n, 0.0);
    for (Int k = 0; k < n; ++k) {
        w[k] = d[ord[k]];
        for (Int i = 0; i < n; ++i) Vout[(size_t)i * n + k] = at(i, ord[k]);
    }
}

/* ------------------------------------------------------------------------------------------------
 *  Ground-state solve standing in for SLEPc EPSSolve (include/DMRGBlockContainer.hpp:1484-1500):
 *  EPS_HEP / EPS_SMALLEST_REAL / nev=1 Krylov-Schur == thick-restart Lanczos with full
 *  re-orthogonalisation; stop when ||r|| <= tol*|theta| (SLEPc default criterion).  Arithmetic of
 *  SLEPc 3.8.3 is not in the tree: only converged (E0, psi up to sign) are comparable.
 * ---------------------------------------------------------------------------------------------- */
struct EigsStats { Int nmatvec = 0, nrestart = 0; Real resid = 0; bool converged = false; };

template <class MatVec>
inline Real LanczosSmallest(Int N, MatVec&& mv, std::vector<Real>& psi, Real tol = 1e-12, Int ncv_in = 16, Int max_it = 1000,
                            EigsStats* st = nullptr, uint64_t seed = 20261018ULL) {
    EigsStats stats;
    if (N == 1) {
        std::vector<Real> x(1, 1.0), y(1);
        mv(x.data(), y.data());
        psi = {1.0};
        stats.nmatvec = 1; stats.converged = true;
        if (st) *st = stats;
        return y[0];
    }
    const Int ld = std::max<Int>(2, std::min(ncv_in, N));
    Int nc = ld; /* current basis size (shrinks only on an invariant subspace) */
    std::vector<std::vector<Real>> V(ld + 1, std::vector<Real>(N));
    std::vector<Real> T((size_t)ld * ld, 0.0); /* T = V^T H V, filled from the orthogonalisation coefficients */
    /* deterministic start vector (splitmix64) */
    uint64_t s = seed;
    auto rnd = [&]() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= (z >> 31);
        return (Real)(z >> 11) / 9007199254740992.0 - 0.5;
    };
    Real nrm = 0;
    for (Int i = 0; i < N; ++i) { V[0][i] = rnd(); nrm += V[0][i] * V[0][i]; }
    nrm = std::sqrt(nrm);
    for (Int i = 0; i < N; ++i) V[0][i] /= nrm;
    Int k = 0; /* number of Ritz vectors kept at the last restart */
    Real theta = 0, resid = 0;
    std::vector<Real> w(N), h(ld + 1);
    for (Int it = 0; it < max_it; ++it) {
        Real beta_last = 0;
        bool invariant = false;
        for (Int j = k; j < nc; ++j) {
            mv(V[j].data(), w.data());
            stats.nmatvec++;
            /* full orthogonalisation against the whole basis, classical Gram-Schmidt applied twice */
            for (Int i = 0; i <= j; ++i) T[(size_t)i * ld + j] = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                for (Int i = 0; i <= j; ++i) {
                    Real d = 0;
                    for (Int q = 0; q < N; ++q) d += V[i][q] * w[q];
                    h[i] = d;
                }
                for (Int i = 0; i <= j; ++i) {
                    const Real d = h[i];
                    for (Int q = 0; q < N; ++q) w[q] -= d * V[i][q];
                    T[(size_t)i * ld + j] += d;
                }
            }
            for (Int i = 0; i <= j; ++i) T[(size_t)j * ld + i] = T[(size_t)i * ld + j];
            Real b = 0;
            for (Int q = 0; q < N; ++q) b += w[q] * w[q];
            b = std::sqrt(b);
            beta_last = b;
            if (b < 1e-14) { nc = j + 1; invariant = true; break; }
            for (Int q = 0; q < N; ++q) V[j + 1][q] = w[q] / b;
        }
        /* Rayleigh-Ritz on the nc×nc projected matrix */
        std::vector<Real> Tm((size_t)nc * nc);
        for (Int i = 0; i < nc; ++i) for (Int j = 0; j < nc; ++j) Tm[(size_t)i * nc + j] = T[(size_t)i * ld + j];
        std::vector<Real> ev, S;
        SymEigDescending(nc, Tm, ev, S); /* descending: the smallest is last */
        theta = ev[nc - 1];
        resid = invariant ? 0.0 : std::fabs(beta_last * S[(size_t)(nc - 1) * nc + (nc - 1)]);
        const bool conv = resid <= tol * std::max(std::fabs(theta), 1e-300);
        const Int kk = conv ? 1 : std::max<Int>(1, std::min<Int>(nc / 2, nc - 1));
        std::vector<std::vector<Real>> Y(kk, std::vector<Real>(N, 0.0));
        for (Int a = 0; a < kk; ++a) {
            const Int colS = nc - 1 - a;
            for (Int i = 0; i < nc; ++i) {
                const Real c = S[(size_t)i * nc + colS];
                for (Int q = 0; q < N; ++q) Y[a][q] += c * V[i][q];
            }
        }
        psi = Y[0];
        if (conv) { stats.converged = true; break; }
        /* thick restart: kept Ritz vectors + the residual direction; T restarts as diag(theta_a), the
           arrow entries re-appear from the projections at step j = kk */
        std::vector<Real> vnext = V[nc];
        std::fill(T.begin(), T.end(), 0.0);
        for (Int a = 0; a < kk; ++a) { V[a] = Y[a]; T[(size_t)a * ld + a] = ev[nc - 1 - a]; }
        V[kk] = vnext;
        k = kk;
        stats.nrestart++;
    }
    Real n2 = 0;
    for (Real v : psi) n2 += v * v;
    n2 = std::sqrt(n2);
    for (Real& v : psi) v /= n2;
    stats.resid = resid;
    if (st) *st = stats;
    return theta;
}

/* ------------------------------------------------------------------------------------------------
 *  GetTruncation — include/DMRGBlockContainer.hpp:1656-1959 (+ Eigen_t :82-94, EigRDM_BlockDiag
 *  :1962-2003, FillRotation_BlockDiag :2006-2057)
 * ---------------------------------------------------------------------------------------------- */
struct Eigen_t { Real eigval; Int seqIdx, epsIdx, blkIdx; };

struct BasisTransformation {
    CSR RotMatT;                 /* m × NStates, rows = kept eigenvectors */
    QuantumNumbers QN;           /* new sector list */
    Real TruncErr = 0;
    std::vector<Eigen_t> spectrum; /* unsorted (grouped) eigenvalues, as dumped by SaveEntanglementSpectra */
    bool tie_at_cut = false;     /* oracle-only diagnostic: |λ_m − λ_{m+1}| <= 1e-13·λ_1 */
};

inline void GetTruncationSide(const KronBlocks_t& KB, const Real* v, Int MStates, bool left, BasisTransformation& BT) {
    const Block& Blk = left ? KB.LeftBlock : KB.RightBlock;
    std::vector<Eigen_t> eigen;
    std::vector<std::vector<Real>> evecs; /* per seqIdx: n×n row-major, eigenvector k in column k */
    std::vector<Int> dims;
    for (Int idx = 0; idx < KB.size(); ++idx) {
        const Int Istart = KB.Offsets(idx);
        const Int Idx_L = KB.LeftIdx(idx), Idx_R = KB.RightIdx(idx);
        const Int N_L = KB.LeftBlock.Magnetization.Sizes(Idx_L);
        const Int N_R = KB.RightBlock.Magnetization.Sizes(Idx_R);
        if (KB.Offsets(idx + 1) - Istart != N_L * N_R) ORACLE_THROW(1, "Incorrect segment length.");
        /* Psi[l][r] = v[Istart + l*N_R + r]  (PsiT is the column-major N_R×N_L view, :1731) */
        const Real* Psi = v + Istart;
        const Int n = left ? N_L : N_R;
        std::vector<Real> rdm((size_t)n * n, 0.0);
        if (left) { /* rdmd_L = Psi PsiT (:1733) */
            for (Int a = 0; a < N_L; ++a)
                for (Int b = a; b < N_L; ++b) {
                    Real s = 0;
                    for (Int r = 0; r < N_R; ++r) s += Psi[a * N_R + r] * Psi[b * N_R + r];
                    rdm[(size_t)a * n + b] = s; rdm[(size_t)b * n + a] = s;
                }
        } else { /* rdmd_R = PsiT Psi (:1734) */
            for (Int l = 0; l < N_L; ++l)
                for (Int a = 0; a < N_R; ++a) {
                    Real pa = Psi[l * N_R + a];
                    if (pa == 0.0) continue;
                    for (Int b = 0; b < N_R; ++b) rdm[(size_t)a * n + b] += pa * Psi[l * N_R + b];
                }
        }
        std::vector<Real> w, V;
        SymEigDescending(n, rdm, w, V);
        for (Int k = 0; k < n; ++k) eigen.push_back({w[k], idx, k, left ? Idx_L : Idx_R});
        evecs.push_back(std::move(V));
        dims.push_back(n);
    }
    BT.spectrum = eigen;
    std::stable_sort(eigen.begin(), eigen.end(), [](const Eigen_t& a, const Eigen_t& b) { return a.eigval > b.eigval; });
    const Int NEig = (Int)eigen.size();
    const Int m = std::min(MStates, NEig);
    BT.tie_at_cut = (m < NEig) && (std::fabs(eigen[m - 1].eigval - eigen[m].eigval) <= 1e-13 * std::fabs(eigen[0].eigval));
    eigen.resize(m);
    std::stable_sort(eigen.begin(), eigen.end(), [](const Eigen_t& a, const Eigen_t& b) { return a.blkIdx < b.blkIdx; });
    const Int NStates = Blk.Magnetization.NumStates();
    BT.RotMatT = CSR(m, NStates);
    Int rowCtr = 0;
    for (const Eigen_t& eig : eigen) {
        const Int startIdx = Blk.Magnetization.Offsets(eig.blkIdx);
        const Int numStates = Blk.Magnetization.Sizes(eig.blkIdx);
        const std::vector<Real>& V = evecs[eig.seqIdx];
        const Int n = dims[eig.seqIdx];
        for (Int i = 0; i < numStates; ++i) { BT.RotMatT.col.push_back(startIdx + i); BT.RotMatT.val.push_back(V[(size_t)i * n + eig.epsIdx]); }
        BT.RotMatT.rowptr[++rowCtr] = (Int)BT.RotMatT.col.size();
    }
    BT.TruncErr = 1.0;
    for (const Eigen_t& eig : eigen) BT.TruncErr -= (eig.eigval > 0) * eig.eigval;
    std::map<Int, Int> BlockIdxs;
    for (const Eigen_t& eig : eigen) BlockIdxs[eig.blkIdx] += 1;
    std::vector<Real> qn_list;
    std::vector<Int> qn_size;
    for (auto& kv : BlockIdxs) { qn_list.push_back(Blk.Magnetization.List(kv.first)); qn_size.push_back(kv.second); }
    BT.QN.Initialize(qn_list, qn_size);
}

inline void GetTruncation(const KronBlocks_t& KB, const std::vector<Real>& gsv, Int MStates, BasisTransformation& BT_L,
                          BasisTransformation& BT_R) {
    if ((Int)gsv.size() != KB.NumStates()) ORACLE_THROW(1, "Incorrect vector length.");
    GetTruncationSide(KB, gsv.data(), MStates, true, BT_L);
    GetTruncationSide(KB, gsv.data(), MStates, false, BT_R);
}

/* ------------------------------------------------------------------------------------------------
 *  RotateOperators — src/DMRGBlock.cpp:677-823:  O' = RotMatT · O · RotMatTᴴ  (MatMatMatMult)
 * ---------------------------------------------------------------------------------------------- */
inline CSR MatMatMatMult_RORt(const CSR& R, const CSR& O, const CSR& Rt /* = R^T */) {
    const Int m = R.nrows, N = O.ncols;
    CSR out(m, m);
    std::vector<Real> acc(N, 0.0), acc2(m, 0.0);
    std::vector<char> mark(N, 0), mark2(m, 0);
    std::vector<Int> list, list2;
    for (Int k = 0; k < m; ++k) {
        list.clear();
        for (Int a = R.rowptr[k]; a < R.rowptr[k + 1]; ++a) {
            const Int c = R.col[a];
            const Real v = R.val[a];
            for (Int b = O.rowptr[c]; b < O.rowptr[c + 1]; ++b) {
                const Int j = O.col[b];
                if (!mark[j]) { mark[j] = 1; list.push_back(j); }
                acc[j] += v * O.val[b];
            }
        }
        list2.clear();
        for (Int j : list) {
            const Real t = acc[j];
            for (Int a = Rt.rowptr[j]; a < Rt.rowptr[j + 1]; ++a) {
                const Int k2 = Rt.col[a];
                if (!mark2[k2]) { mark2[k2] = 1; list2.push_back(k2); }
                acc2[k2] += t * Rt.val[a];
            }
            acc[j] = 0.0; mark[j] = 0;
        }
        std::sort(list2.begin(), list2.end());
        for (Int k2 : list2) { out.col.push_back(k2); out.val.push_back(acc2[k2]); acc2[k2] = 0.0; mark2[k2] = 0; }
        out.rowptr[k + 1] = (Int)out.col.size();
    }
    return out;
}

inline void RotateOperators(Block& Dest, const Block& Source, const CSR& RotMatT) {
    if (RotMatT.ncols != Source.NumStates()) ORACLE_THROW(1, "RotMatT_in incorrect number of cols.");
    if (RotMatT.nrows != Dest.NumStates()) ORACLE_THROW(1, "RotMatT_in incorrect number of rows.");
    if (Source.NumSites() != Dest.NumSites()) ORACLE_THROW(1, "RotMatT_in incorrect number of sites.");
    CSR RotMat = RotMatT.Transpose();
    for (Int i = 0; i < Dest.num_sites; ++i) {
        Dest.SpData[i] = MatMatMatMult_RORt(RotMatT, Source.SpData[i], RotMat);
        Dest.SzData[i] = MatMatMatMult_RORt(RotMatT, Source.SzData[i], RotMat);
    }
    Dest.H = MatMatMatMult_RORt(RotMatT, Source.H, RotMat);
    Dest.CheckOperatorBlocks();
}

/* ------------------------------------------------------------------------------------------------
 *  DMRG container — include/DMRGBlockContainer.hpp: Warmup :687-861, Sweeps :864-993,
 *  SingleSweep :996-1088, SingleDMRGStep :1304-1653.  Disk scratch, JSON and correlators omitted.
 * ---------------------------------------------------------------------------------------------- */
struct StepData {
    Int GlobIdx, LoopType /*0 warmup,1 sweep*/, LoopIdx, StepIdx;
    Int NumSites_Sys, NumSites_Env, NumSites_SysEnl, NumSites_EnvEnl;
    Int NumStates_Sys, NumStates_Env, NumStates_SysEnl, NumStates_EnvEnl, NumStates_SysRot, NumStates_EnvRot, NumStates_H;
    Real TruncErr_Sys, TruncErr_Env, GSEnergy;
    /* oracle extras for parity tests */
    std::vector<Real> qn_list_L, qn_list_R;
    std::vector<Int> qn_size_L, qn_size_R;
    Int nmatvec; bool tie_L, tie_R;
};

struct DMRG {
    Hamiltonian Ham;
    Block AddSite;
    std::vector<Block> sys_blocks;
    Int num_sites = 0, sys_ninit = 0, mwarmup = 0;
    Real qn_sector = 0.0; /* :1176 */
    Real eps_tol = 1e-12; Int eps_ncv = 16, eps_max_it = 2000;
    Int GlobIdx = 0, LoopIdx = 0, StepIdx = 0, LoopType = 0;
    Real gse = 0;
    std::vector<Real> trunc_err;
    std::vector<StepData> steps;
    bool warmed_up = false;
    int spin_twice = 1;
    /* optional hook: called with the shell and enlarged blocks of every step (parity-fixture dumps) */
    void (*step_hook)(void*, const DMRG&, const KronBlocks_t&, const KronSumShell&, const std::vector<Real>&, const StepData&) = nullptr;
    void* hook_ctx = nullptr;

    void Initialize() {
        AddSite.InitializeSingleSite(spin_twice);
        num_sites = Ham.NumSites();
        if (num_sites < 2) ORACLE_THROW(1, "There must be at least two total sites.");
        if (num_sites % 2) ORACLE_THROW(1, "Total number of sites must be even.");
    }

    void SingleDMRGStep(Block& SysBlock, Block& EnvBlock, Int MStates, Block& SysBlockOut, Block& EnvBlockOut) {
        StepData sd{};
        sd.NumSites_Sys = SysBlock.NumSites(); sd.NumSites_Env = EnvBlock.NumSites();
        sd.NumStates_Sys = SysBlock.NumStates(); sd.NumStates_Env = EnvBlock.NumStates();
        const bool flg = (&SysBlock == &EnvBlock);
        Block SysBlockEnl, EnvBlockEnlStore;
        KronEye_Explicit(SysBlock, AddSite, Ham.H(SysBlock.NumSites() + 1), SysBlockEnl);
        if (!flg) KronEye_Explicit(EnvBlock, AddSite, Ham.H(EnvBlock.NumSites() + 1), EnvBlockEnlStore);
        Block& EnvBlockEnl = flg ? SysBlockEnl : EnvBlockEnlStore;
        sd.NumSites_SysEnl = SysBlockEnl.NumSites(); sd.NumSites_EnvEnl = EnvBlockEnl.NumSites();
        sd.NumStates_SysEnl = SysBlockEnl.NumStates(); sd.NumStates_EnvEnl = EnvBlockEnl.NumStates();
        const Int NumSitesTotal = SysBlockEnl.NumSites() + EnvBlockEnl.NumSites();
        const std::vector<Term> Terms = Ham.H(NumSitesTotal);
        KronBlocks_t KB(SysBlockEnl, EnvBlockEnl, {qn_sector});
        sd.NumStates_H = KB.NumStates();
        if (KB.NumStates() == 0) ORACLE_THROW(1, "empty target sector");
        KronSumShell shell;
        KB.KronSumConstruct(SysBlockEnl, EnvBlockEnl, Terms, nullptr, &shell);
        std::vector<Real> gsv;
        EigsStats st;
        Real gse_r = LanczosSmallest(KB.NumStates(), [&](const Real* x, Real* y) { shell.MatMult(x, y); }, gsv, eps_tol,
                                     eps_ncv, eps_max_it, &st);
        sd.GSEnergy = gse_r;
        sd.nmatvec = st.nmatvec;
        BasisTransformation BT_L, BT_R;
        GetTruncation(KB, gsv, MStates, BT_L, BT_R);
        sd.GlobIdx = GlobIdx; sd.LoopType = LoopType; sd.LoopIdx = LoopIdx; sd.StepIdx = StepIdx;
        sd.TruncErr_Sys = BT_L.TruncErr; sd.TruncErr_Env = BT_R.TruncErr;
        sd.qn_list_L = BT_L.QN.List(); sd.qn_size_L = BT_L.QN.Sizes();
        sd.qn_list_R = BT_R.QN.List(); sd.qn_size_R = BT_R.QN.Sizes();
        sd.tie_L = BT_L.tie_at_cut; sd.tie_R = BT_R.tie_at_cut;
        if (step_hook) step_hook(hook_ctx, *this, KB, shell, gsv, sd);
        if (SysBlockEnl.init_Sm) SysBlockEnl.DestroySm();
        if (EnvBlockEnl.init_Sm) EnvBlockEnl.DestroySm();
        Block NewSys, NewEnv;
        NewSys.Initialize(SysBlockEnl.NumSites(), BT_L.QN.List(), BT_L.QN.Sizes());
        NewSys.spin_twice = spin_twice;
        RotateOperators(NewSys, SysBlockEnl, BT_L.RotMatT);
        if (!flg) {
            NewEnv.Initialize(EnvBlockEnl.NumSites(), BT_R.QN.List(), BT_R.QN.Sizes());
            NewEnv.spin_twice = spin_twice;
            RotateOperators(NewEnv, EnvBlockEnl, BT_R.RotMatT);
        }
        /* outputs may alias the inputs (sys_blocks[] entries): assign last */
        SysBlockOut = std::move(NewSys);
        if (!flg) EnvBlockOut = std::move(NewEnv);
        sd.NumStates_SysRot = SysBlockOut.NumStates();
        sd.NumStates_EnvRot = EnvBlockOut.NumStates();
        gse = gse_r;
        trunc_err.push_back(BT_L.TruncErr);
        steps.push_back(sd);
        ++GlobIdx; ++StepIdx;
    }

    /* :687-861 */
    void Warmup() {
        if (mwarmup == 0) return;
        sys_blocks.assign(num_sites - 1, Block());
        sys_ninit = 0;
        sys_blocks[sys_ninit++].InitializeSingleSite(spin_twice);
        Int nsites_cluster = Ham.NumEnvSites();
        if (nsites_cluster % 2) nsites_cluster *= 2;
        while (sys_ninit < nsites_cluster) {
            Int NumSitesTotal = sys_blocks[sys_ninit - 1].NumSites() + AddSite.NumSites();
            KronEye_Explicit(sys_blocks[sys_ninit - 1], AddSite, Ham.H(NumSitesTotal), sys_blocks[sys_ninit]);
            ++sys_ninit;
        }
        if (sys_ninit >= num_sites / 2) ORACLE_THROW(1, "No DMRG Steps were performed since all site operators were created exactly.");
        LoopType = 0; StepIdx = 0;
        while (sys_ninit < num_sites / 2) {
            Int full_cluster = (((sys_ninit + 2) / nsites_cluster) + 1) * nsites_cluster;
            Int env_numsites = full_cluster - sys_ninit - 2;
            Int env_add = ((sys_ninit - env_numsites) / nsites_cluster) * nsites_cluster;
            env_numsites += env_add;
            full_cluster += env_add;
            if (env_numsites < 1 || env_numsites > sys_ninit) ORACLE_THROW(1, "Incorrect number of sites.");
            SingleDMRGStep(sys_blocks[sys_ninit - 1], sys_blocks[env_numsites - 1], mwarmup, sys_blocks[sys_ninit],
                           sys_blocks[env_numsites]);
            ++sys_ninit;
        }
        warmed_up = true;
        ++LoopIdx;
    }
    /* :996-1088 */
    void SingleSweep(Int MStates, Int MinBlock = -1) {
        if (!warmed_up) ORACLE_THROW(1, "Warmup must be called first before performing sweeps.");
        trunc_err.clear();
        Int min_block = MinBlock < 0 ? 1 : MinBlock;
        LoopType = 1; StepIdx = 0;
        for (Int iblock = num_sites / 2; iblock < num_sites - min_block - 2; ++iblock) {
            const Int insys = iblock - 1, inenv = num_sites - iblock - 3;
            const Int outsys = iblock, outenv = num_sites - iblock - 2;
            SingleDMRGStep(sys_blocks[insys], sys_blocks[inenv], MStates, sys_blocks[outsys], sys_blocks[outenv]);
        }
        for (Int iblock = min_block; iblock < num_sites / 2; ++iblock) {
            const Int insys = num_sites - iblock - 3, inenv = iblock - 1;
            const Int outsys = num_sites - iblock - 2, outenv = iblock;
            SingleDMRGStep(sys_blocks[insys], sys_blocks[inenv], MStates, sys_blocks[outsys], sys_blocks[outenv]);
        }
        ++LoopIdx;
    }
};

} /* namespace oracle */
