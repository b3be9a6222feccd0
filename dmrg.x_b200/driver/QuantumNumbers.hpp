/*  QuantumNumbers — host mirror of include/QuantumNumbers.hpp:30-239 (pure integer bookkeeping). */
#pragma once
#include <algorithm>
#include <vector>

#include "PetscShim.hpp"

class QuantumNumbers {
public:
    /** Initialize(comm, qn_list, qn_size): src/QuantumNumbers.cpp:9-52 */
    PetscErrorCode Initialize(const MPI_Comm&, const std::vector<PetscReal>& qn_list_in, const std::vector<PetscInt>& qn_size_in) {
        if (qn_list_in.empty()) SETERRQ(0, PETSC_ERR_ARG_WRONG, "Initialization error: Empty input list.");
        if (qn_list_in.size() != qn_size_in.size()) SETERRQ(0, PETSC_ERR_ARG_WRONG, "Initialization error: Input list sizes mismatch.");
        for (size_t i = 1; i < qn_list_in.size(); ++i)
            if (qn_list_in[i] >= qn_list_in[i - 1]) SETERRQ(0, 1, "qn_list_in must be sorted descending.");
        qn_list = qn_list_in; qn_size = qn_size_in;
        qn_offset.assign(qn_list.size() + 1, 0);
        for (size_t i = 0; i < qn_list.size(); ++i) qn_offset[i + 1] = qn_offset[i] + qn_size[i];
        initialized = PETSC_TRUE;
        return 0;
    }
    PetscErrorCode CheckInitialized() const { return initialized ? 0 : PETSC_ERR_ARG_WRONGSTATE; }
    PetscBool Initialized() const { return initialized; }
    PetscInt NumSectors() const { return (PetscInt)qn_list.size(); }
    PetscInt NumStates() const { return qn_offset.empty() ? 0 : qn_offset.back(); }
    const std::vector<PetscReal>& List() const { return qn_list; }
    const std::vector<PetscReal>& ListRef() const { return qn_list; }
    const std::vector<PetscInt>& Sizes() const { return qn_size; }
    const std::vector<PetscInt>& Offsets() const { return qn_offset; }
    /* out-of-range lookups return -1 (include/QuantumNumbers.hpp:100-141) */
    PetscReal List(PetscInt i) const { return qn_list.at((size_t)i); }
    PetscInt Sizes(PetscInt i) const { return (i < 0 || i >= NumSectors()) ? -1 : qn_size[(size_t)i]; }
    PetscInt Offsets(PetscInt i) const { return (i < 0 || i >= NumSectors()) ? -1 : qn_offset[(size_t)i]; }
    /** src/QuantumNumbers.cpp:72-96 */
    PetscErrorCode OpBlockToGlobalRange(PetscInt BlockIdx, PetscInt BlockShift, PetscInt& s, PetscInt& e, PetscBool& flg) const {
        if (BlockIdx < 0 || BlockIdx >= NumSectors()) return PETSC_ERR_ARG_OUTOFRANGE;
        const PetscInt o = BlockIdx + BlockShift;
        if (o < 0 || o >= NumSectors()) { flg = PETSC_FALSE; return 0; }
        flg = PETSC_TRUE; s = qn_offset[(size_t)o]; e = qn_offset[(size_t)o + 1];
        return 0;
    }
    PetscInt OpBlockToGlobalRangeStart(PetscInt BlockIdx, PetscInt BlockShift, PetscBool& flg) const {
        PetscInt s = -1, e = -1;
        if (OpBlockToGlobalRange(BlockIdx, BlockShift, s, e, flg)) return -1;
        return flg ? s : -1;
    }
    /** src/QuantumNumbers.cpp:124-158 */
    PetscErrorCode GlobalIdxToBlockIdx(PetscInt GlobIdx, PetscInt& BlockIdx) const {
        if (GlobIdx < 0 || GlobIdx >= NumStates()) return PETSC_ERR_ARG_OUTOFRANGE;
        BlockIdx = (PetscInt)(std::upper_bound(qn_offset.begin(), qn_offset.end(), GlobIdx) - qn_offset.begin()) - 1;
        return 0;
    }
    PetscInt BlockIdxToGlobalIdx(PetscInt BlockIdx, PetscInt LocIdx) const { return qn_offset[(size_t)BlockIdx] + LocIdx; }

private:
    PetscBool initialized = PETSC_FALSE;
    std::vector<PetscReal> qn_list;
    std::vector<PetscInt> qn_size, qn_offset;
};
