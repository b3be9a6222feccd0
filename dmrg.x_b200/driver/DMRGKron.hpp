/*  KronBlocks_t and the shell Hamiltonian — host mirror of include/DMRGKron.hpp:117-480 over the C ABI. */
#pragma once
#include <memory>
#include <tuple>
#include <vector>

#include "DMRGBlock.hpp"

typedef std::tuple<PetscReal, PetscInt, PetscInt, PetscInt> KronBlock_t; /* include/DMRGKron.hpp:22 */

/** The MATSHELL of the reference (MatCreateShell + MATOP_MULT = MatMult_KronSumShell, src/DMRGKron.cpp:1912-1914) */
struct ShellMat {
    dmrgx_hshell h = nullptr;
    PetscInt N = 0;
    explicit operator bool() const { return h != nullptr; }
};
/** a device vector (Vec) */
struct Vec {
    double* d = nullptr;
    PetscInt n = 0;
};
inline PetscErrorCode MatCreateVecs(const ShellMat& H, Vec* v) { v->n = H.N; return dmrgx_vec_alloc(DmrgxContext(), H.N, &v->d); }
inline PetscErrorCode VecDestroy(Vec* v) { PetscErrorCode e = v->d ? dmrgx_vec_free(DmrgxContext(), v->d) : 0; v->d = nullptr; return e; }
/** include/DMRGKron.hpp:24, src/DMRGKron.cpp:1827-1869 */
inline PetscErrorCode MatMult_KronSumShell(ShellMat A, Vec x, Vec y) { return dmrgx_hshell_apply(A.h, x.d, y.d); }
/** src/DMRGKron.cpp:1919-1942 — the caller destroys the shell explicitly (include/DMRGBlockContainer.hpp:1502-1505) */
inline PetscErrorCode MatDestroy_KronSumShell(ShellMat* p_mat) { PetscErrorCode e = p_mat->h ? dmrgx_hshell_destroy(p_mat->h) : 0; p_mat->h = nullptr; return e; }

class KronBlocks_t {
public:
    /** include/DMRGKron.hpp:124-213 — constructors throw std::runtime_error (:137-144) */
    KronBlocks_t(Block::SpinBase& LeftBlock, Block::SpinBase& RightBlock, const std::vector<PetscReal>& QNSectors, FILE* fp_prealloc,
                 const PetscInt& GlobIdx)
        : GlobIdx(GlobIdx), LeftBlock(LeftBlock), RightBlock(RightBlock), fp_prealloc(fp_prealloc) {
        if (!LeftBlock.Initialized()) throw std::runtime_error("Left input block not initialized.");
        if (!RightBlock.Initialized()) throw std::runtime_error("Right input block not initialized.");
        dmrgx_kron k;
        if (dmrgx_kron_create(LeftBlock.Handle(), RightBlock.Handle(), (dmrgx_int)QNSectors.size(), QNSectors.data(), &k))
            throw std::runtime_error(dmrgx_last_error());
        h = std::shared_ptr<dmrgx_kron_s>(k, [](dmrgx_kron p) { if (p) dmrgx_kron_destroy(p); });
        num_blocks = dmrgx_kron_size(k);
        kb_list.resize((size_t)num_blocks); kb_size.resize((size_t)num_blocks); kb_offset.resize((size_t)num_blocks + 1);
        std::vector<PetscInt> li((size_t)num_blocks), ri((size_t)num_blocks);
        dmrgx_kron_data(k, kb_list.data(), li.data(), ri.data(), kb_size.data(), kb_offset.data());
        for (PetscInt i = 0; i < num_blocks; ++i) KronBlocks.push_back(std::make_tuple(kb_list[i], li[i], ri[i], kb_size[i]));
        num_states = dmrgx_kron_num_states(k);
    }
    PetscInt size() const { return (PetscInt)KronBlocks.size(); }
    const std::vector<KronBlock_t>& data() const { return KronBlocks; }
    KronBlock_t data(size_t idx) const { return KronBlocks[idx]; }
    KronBlock_t operator[](size_t idx) const { return KronBlocks[idx]; }
    std::vector<PetscReal> List() const { return kb_list; }
    std::vector<PetscInt> Offsets() const { return kb_offset; }
    PetscInt Offsets(const PetscInt& idx) const { return kb_offset[(size_t)idx]; }
    PetscReal QN(const PetscInt& idx) const { return std::get<0>(KronBlocks[(size_t)idx]); }
    PetscInt LeftIdx(const PetscInt& idx) const { return std::get<1>(KronBlocks[(size_t)idx]); }
    PetscInt RightIdx(const PetscInt& idx) const { return std::get<2>(KronBlocks[(size_t)idx]); }
    PetscInt Sizes(const PetscInt& idx) const { return std::get<3>(KronBlocks[(size_t)idx]); }
    std::vector<PetscInt> Sizes() const { return kb_size; }
    const Block::SpinBase& LeftBlockRef() const { return LeftBlock; }
    const Block::SpinBase& RightBlockRef() const { return RightBlock; }
    Block::SpinBase& LeftBlockRefMod() { return LeftBlock; }
    Block::SpinBase& RightBlockRefMod() { return RightBlock; }
    /** :272-294: -1 when the pair is absent */
    PetscInt Offsets(const PetscInt& lidx, const PetscInt& ridx) const { return dmrgx_kron_offsets_lr(h.get(), lidx, ridx); }
    PetscInt Map(const PetscInt& lidx, const PetscInt& ridx) const { return dmrgx_kron_map(h.get(), lidx, ridx); }
    PetscInt NumStates() const { return num_states; }
    dmrgx_kron Handle() const { return h.get(); }

    /** KronSumConstruct(Terms, MatOut): include/DMRGKron.hpp:300, src/DMRGKron.cpp:759-841.  Only the shell form exists here
        (the reference's default in SingleDMRGStep, include/DMRGBlockContainer.hpp:1252). */
    PetscErrorCode KronSumConstruct(const std::vector<Hamiltonians::Term>& Terms, ShellMat& MatOut) {
        if (!do_shell) SETERRQ(0, PETSC_ERR_SUP, "explicit superblock matrices are not built on the device path (use -do_shell 1)");
        std::vector<double> a; std::vector<int> iop, jop; std::vector<dmrgx_int> is, js;
        for (const auto& t : Terms) { a.push_back(t.a); iop.push_back(t.Iop); is.push_back(t.Isite); jop.push_back(t.Jop); js.push_back(t.Jsite); }
        DMRGX_CALL(dmrgx_hshell_create(h.get(), (dmrgx_int)Terms.size(), a.data(), iop.data(), is.data(), jop.data(), js.data(), &MatOut.h));
        MatOut.N = num_states;
        return 0;
    }
    /** KronConstruct(Mat_L, OpType_L, Mat_R, OpType_R, MatOut): include/DMRGKron.hpp:309 */
    PetscErrorCode KronConstruct(const Mat& Mat_L, const Op_t& OpType_L, const Mat& Mat_R, const Op_t& OpType_R, ShellMat& MatOut) {
        DMRGX_CALL(dmrgx_hshell_create_single(h.get(), Mat_L ? (int)OpType_L : DMRGX_OP_EYE, Mat_L.isite, Mat_R ? (int)OpType_R : DMRGX_OP_EYE,
                                              Mat_R.isite, &MatOut.h));
        MatOut.N = num_states;
        return 0;
    }
    PetscErrorCode KronSumSetShellMatrix(const PetscBool& do_shell_in) { do_shell = do_shell_in; return 0; }     /* :318 */
    PetscErrorCode KronSumSetRedistribute(const PetscBool& do_redistribute_in = PETSC_TRUE) { do_redistribute = do_redistribute_in; return 0; } /* :324 */
    /** :332-337.  In the reference -ks_tol reaches only the KronBlocks_t of the SUPERBLOCK (include/DMRGBlockContainer.hpp:1449)
        and is read only by the explicit MPIAIJ construction (src/DMRGKron.cpp:1109-1112, 1449-1454), never by the shell
        matvec; the KronBlocks_t inside KronEye_Explicit (src/DMRGKron.cpp:612) keeps the default 1e-16, which is what the
        device enlargement applies (csrc/block.cpp).  With the shell form (the only one built here) the option therefore has
        no effect on either side: it is parsed for compatibility and deliberately not forwarded. */
    PetscErrorCode KronSumSetToleranceFromOptions() { return PetscOptions::DB().GetReal("-ks_tol", &ks_tol, NULL); }

private:
    PetscInt GlobIdx;
    Block::SpinBase& LeftBlock;
    Block::SpinBase& RightBlock;
    FILE* fp_prealloc;
    std::shared_ptr<dmrgx_kron_s> h;
    std::vector<KronBlock_t> KronBlocks;
    std::vector<PetscReal> kb_list;
    std::vector<PetscInt> kb_size, kb_offset;
    PetscInt num_blocks = 0, num_states = 0;
    PetscBool do_shell = PETSC_TRUE, do_redistribute = PETSC_FALSE;
    PetscReal ks_tol = 1.0e-16;
};

/** KronEye_Explicit(LeftBlock, RightBlock, Terms, BlockOut): include/DMRGKron.hpp:484, src/DMRGKron.cpp:459-615 */
inline PetscErrorCode KronEye_Explicit(Block::SpinBase& LeftBlock, Block::SpinBase& RightBlock, const std::vector<Hamiltonians::Term>& Terms,
                                       Block::SpinBase& BlockOut) {
    if (!LeftBlock.Initialized()) SETERRQ(0, 1, "Left input block not initialized.");
    if (!RightBlock.Initialized()) SETERRQ(0, 1, "Right input block not initialized.");
    std::vector<double> a; std::vector<int> iop, jop; std::vector<dmrgx_int> is, js;
    for (const auto& t : Terms) { a.push_back(t.a); iop.push_back(t.Iop); is.push_back(t.Isite); jop.push_back(t.Jop); js.push_back(t.Jsite); }
    dmrgx_block out;
    DMRGX_CALL(dmrgx_block_enlarge(LeftBlock.Handle(), RightBlock.Handle(), (dmrgx_int)Terms.size(), a.data(), iop.data(), is.data(), jop.data(),
                                   js.data(), &out));
    return BlockOut.Adopt(out);
}
