/*  Block::SpinBase — host mirror of include/DMRGBlock.hpp:79-434 over the device block of the C ABI.
 *  The operator matrices live in HBM as tiles; `Mat` here is a light (block, operator, site) reference.
 *  Copying a SpinBase shares the device block, like copying the reference object shares its Mat handles
 *  (tests/UnitTests_DMRGBlock.cpp:56-70). */
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include "../../include/dmrgx.h"
#include "Hamiltonians.hpp"
#include "PetscShim.hpp"
#include "QuantumNumbers.hpp"

/** The device context all blocks of this process live in (stands in for the MPI communicator). */
inline dmrgx_ctx& DmrgxContext() { static dmrgx_ctx ctx = nullptr; return ctx; }

#define DMRGX_CALL(expr) do { int e_ = (expr); if (e_) { fprintf(stderr, "[dmrgx] %s -> %d: %s\n", #expr, e_, dmrgx_last_error()); return e_; } } while (0)

/** A reference to one operator matrix of a device block (what `Mat` is in the reference's signatures). */
struct Mat {
    dmrgx_block blk = nullptr;
    int op = 0;
    PetscInt isite = 0;
    explicit operator bool() const { return blk != nullptr; }
};

namespace Block {

class SpinBase {
public:
    /** Initialize(comm): include/DMRGBlock.hpp:226 */
    PetscErrorCode Initialize(const MPI_Comm&) { mpi_init = PETSC_TRUE; return 0; }
    /** Initialize(comm, num_sites, num_states): src/DMRGBlock.cpp:44-170 — only the single-site form creates operators */
    PetscErrorCode Initialize(const MPI_Comm&, const PetscInt& num_sites_in, const PetscInt& num_states_in, const PetscBool& init_ops = PETSC_TRUE) {
        if (!(num_sites_in == 1 && num_states_in == PETSC_DEFAULT && init_ops))
            SETERRQ(0, PETSC_ERR_SUP, "only Initialize(comm, 1, PETSC_DEFAULT) or the sector-list form is supported");
        std::string spin; PetscBool set;
        PetscOptions::DB().GetString("-spin", spin, &set);
        int spin_twice = 1;
        if (set) { if (spin == "1/2") spin_twice = 1; else if (spin == "1") spin_twice = 2; else SETERRQ1(0, 1, "Given -spin %s not valid/implemented.", spin.c_str()); }
        dmrgx_block b;
        DMRGX_CALL(dmrgx_block_single_site(DmrgxContext(), spin_twice, &b));
        return Adopt(b);
    }
    /** Initialize(comm, num_sites, qn_list, qn_size): src/DMRGBlock.cpp:173-196 */
    PetscErrorCode Initialize(const MPI_Comm&, const PetscInt& num_sites_in, const std::vector<PetscReal>& qn_list_in,
                              const std::vector<PetscInt>& qn_size_in) {
        dmrgx_block b;
        DMRGX_CALL(dmrgx_block_create(DmrgxContext(), num_sites_in, (dmrgx_int)qn_list_in.size(), qn_list_in.data(), qn_size_in.data(), &b));
        return Adopt(b);
    }
    /** take ownership of a device block produced by the library (enlargement, rotation) */
    PetscErrorCode Adopt(dmrgx_block b) {
        h = std::shared_ptr<dmrgx_block_s>(b, [](dmrgx_block p) { if (p) dmrgx_block_destroy(p); });
        static long long next_serial = 0;
        serial = ++next_serial; /* identity of this block's contents (copies share it): -wavefunction_prediction checks that blocks chain up */
        dmrgx_int ns, nst, nsec;
        DMRGX_CALL(dmrgx_block_info(b, &ns, &nst, &nsec));
        num_sites = ns; num_states = nst;
        std::vector<PetscReal> qn((size_t)nsec); std::vector<PetscInt> sz((size_t)nsec);
        DMRGX_CALL(dmrgx_block_sectors(b, qn.data(), sz.data()));
        PetscErrorCode ierr = Magnetization.Initialize(0, qn, sz); CHKERRQ(ierr);
        init = PETSC_TRUE;
        return 0;
    }
    PetscBool Initialized() const { return init; }
    long long Serial() const { return serial; }
    MPI_Comm MPIComm() const { return 0; }
    PetscInt NumSites() const { return num_sites; }
    PetscInt NumStates() const { return num_states; }
    dmrgx_block Handle() const { return h.get(); }
    /** include/DMRGBlock.hpp:350-378: accessors throw on a bad site */
    Mat Sz(const PetscInt& Isite) const { Check(Isite); return Mat{h.get(), DMRGX_OP_SZ, Isite}; }
    Mat Sp(const PetscInt& Isite) const { Check(Isite); return Mat{h.get(), DMRGX_OP_SP, Isite}; }
    Mat Sm(const PetscInt& Isite) const { Check(Isite); return Mat{h.get(), DMRGX_OP_SM, Isite}; }
    Mat H() const { return Mat{h.get(), DMRGX_OP_H, 0}; }
    /** the reference transposes Sp on demand (src/DMRGBlock.cpp:623-646); here Sm is always available as a view */
    PetscErrorCode CreateSm() { return 0; }
    PetscErrorCode DestroySm() { return 0; }
    /** src/DMRGBlock.cpp:413-447, 603-620 */
    PetscErrorCode CheckOperators() const { return init ? 0 : PETSC_ERR_ARG_CORRUPT; }
    PetscErrorCode CheckSectors() const { return Magnetization.NumStates() == num_states ? 0 : PETSC_ERR_ARG_WRONG; }
    PetscErrorCode CheckOperatorBlocks() const { if (!init) return PETSC_ERR_ARG_CORRUPT; return dmrgx_block_check(h.get()); }
    /** set an operator from CSR with global column indices (the layout MatGetRow returns) */
    PetscErrorCode MatSetFromCSR(int op, PetscInt isite, const PetscInt* rowptr, const PetscInt* col, const PetscScalar* val) {
        return dmrgx_block_set_operator(h.get(), op, isite, rowptr, col, val);
    }
    /** blocks stay resident in HBM: the reference's scratch round trips (src/DMRGBlock.cpp:889-1103) become no-ops */
    PetscErrorCode EnsureSaved() { return 0; }
    PetscErrorCode EnsureRetrieved() { return 0; }
    PetscErrorCode InitializeSave(const std::string&) { return 0; }
    /** InitializeFromDisk(comm, block_path): src/DMRGBlock.cpp:214-372 — implemented in DMRGBlockIO.hpp (BlockIO::Load) */
    PetscErrorCode InitializeFromDisk(const MPI_Comm&, const std::string& block_path);
    PetscErrorCode SetDiskStorage(const std::string&, const std::string&) { return 0; }
    /** src/DMRGBlock.cpp:825-887 */
    PetscErrorCode Destroy() { h.reset(); init = PETSC_FALSE; num_sites = num_states = 0; return 0; }

    QuantumNumbers Magnetization;

private:
    void Check(const PetscInt& Isite) const { if (Isite < 0 || Isite >= num_sites) throw std::runtime_error("Attempted to access non-existent site."); }
    std::shared_ptr<dmrgx_block_s> h;
    PetscBool init = PETSC_FALSE, mpi_init = PETSC_FALSE;
    PetscInt num_sites = 0, num_states = 0;
    long long serial = 0;
};

}  // namespace Block
