/*  On-disk block format of the reference (src/DMRGBlock.cpp:889-1103, 214-372): a directory per block holding
 *      BlockInfo.dat        key/value lines: NumBytesPetscInt, NumBytesPetscScalar, PetscUseComplex, SpinTypeKey, NumSites, NumStates, NumSectors
 *      QuantumNumbers.dat   one "size qn" line per sector
 *      Sz_%09d.mat, Sp_%09d.mat (one per site), H_000000000.mat     PETSc binary AIJ matrices (MatView on a binary viewer)
 *  PETSc binary Mat layout [PETSc 3.8 MatView_SeqAIJ_Binary / MatLoad, not in the reference tree]: big-endian
 *      PetscInt  MAT_FILE_CLASSID = 1211216, M, N, nz;  PetscInt rowlen[M];  PetscInt col[nz];  PetscScalar val[nz]
 *  with PetscInt 4 bytes in a default build and 8 with --with-64-bit-indices (BlockInfo.dat says which).
 *  Blocks live in HBM here; this is the host-side import / export through dmrgx_block_get_operator / _set_operator, used by
 *  -restart_dir, by the per-sweep checkpoints and to exchange blocks with the reference (BASELINE configs[4]).
 */
#pragma once
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>

#include "DMRGBlock.hpp"

namespace BlockIO {

constexpr long long MAT_FILE_CLASSID = 1211216;

inline std::string OpFilename(const std::string& RootDir, const std::string& OpName, const size_t& isite = 0) { /* :889-893 */
    std::ostringstream oss;
    oss << RootDir << OpName << "_" << std::setfill('0') << std::setw(9) << isite << ".mat";
    return oss.str();
}

inline void put_be(std::ostream& os, unsigned long long v, int nbytes) {
    unsigned char b[8];
    for (int i = 0; i < nbytes; ++i) b[i] = (unsigned char)(v >> (8 * (nbytes - 1 - i)));
    os.write((const char*)b, nbytes);
}
inline unsigned long long get_be(std::istream& is, int nbytes) {
    unsigned char b[8] = {0};
    is.read((char*)b, nbytes);
    unsigned long long v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | b[i];
    return v;
}
inline long long get_int(std::istream& is, int nbytes) {
    const unsigned long long v = get_be(is, nbytes);
    return nbytes == 4 ? (long long)(int32_t)(uint32_t)v : (long long)v;
}

inline PetscErrorCode WriteMat(const std::string& file, PetscInt n, const std::vector<PetscInt>& rowptr, const std::vector<PetscInt>& col,
                               const std::vector<PetscScalar>& val, int int_bytes) {
    std::ofstream os(file.c_str(), std::ios::binary);
    if (!os) SETERRQ1(0, 1, "cannot write %s", file.c_str());
    const PetscInt nz = rowptr[(size_t)n];
    put_be(os, (unsigned long long)MAT_FILE_CLASSID, int_bytes);
    put_be(os, (unsigned long long)n, int_bytes); put_be(os, (unsigned long long)n, int_bytes); put_be(os, (unsigned long long)nz, int_bytes);
    for (PetscInt r = 0; r < n; ++r) put_be(os, (unsigned long long)(rowptr[(size_t)r + 1] - rowptr[(size_t)r]), int_bytes);
    for (PetscInt e = 0; e < nz; ++e) put_be(os, (unsigned long long)col[(size_t)e], int_bytes);
    for (PetscInt e = 0; e < nz; ++e) { unsigned long long bits; std::memcpy(&bits, &val[(size_t)e], 8); put_be(os, bits, 8); }
    return os.good() ? 0 : 1;
}

inline PetscErrorCode ReadMat(const std::string& file, PetscInt n_expected, std::vector<PetscInt>& rowptr, std::vector<PetscInt>& col,
                              std::vector<PetscScalar>& val, int int_bytes) {
    std::ifstream is(file.c_str(), std::ios::binary);
    if (!is) SETERRQ1(0, 1, "cannot read %s", file.c_str());
    if (get_int(is, int_bytes) != MAT_FILE_CLASSID) SETERRQ1(0, PETSC_ERR_ARG_WRONG, "%s is not a PETSc binary matrix (or was written with another PetscInt size)", file.c_str());
    const long long M = get_int(is, int_bytes), N = get_int(is, int_bytes), nz = get_int(is, int_bytes);
    if (M != n_expected || N != n_expected || nz < 0) SETERRQ1(0, PETSC_ERR_ARG_WRONG, "%s: matrix dimensions do not match the block", file.c_str());
    { /* nz must fit in the file: header + M row lengths + nz columns + nz values */
        is.seekg(0, std::ios::end);
        const long long fsize = (long long)is.tellg();
        is.seekg(4LL * int_bytes, std::ios::beg);
        if (fsize < 4LL * int_bytes + M * int_bytes || (fsize - 4LL * int_bytes - M * int_bytes) / (int_bytes + 8) < nz)
            SETERRQ1(0, PETSC_ERR_ARG_CORRUPT, "%s: file is shorter than its header says", file.c_str());
    }
    rowptr.assign((size_t)M + 1, 0);
    for (long long r = 0; r < M; ++r) {
        const long long len = get_int(is, int_bytes);
        if (len < 0 || len > nz) SETERRQ1(0, PETSC_ERR_ARG_CORRUPT, "%s: invalid row length", file.c_str());
        rowptr[(size_t)r + 1] = rowptr[(size_t)r] + len;
    }
    if (rowptr[(size_t)M] != nz) SETERRQ1(0, PETSC_ERR_ARG_CORRUPT, "%s: row lengths do not add up to nz", file.c_str());
    col.resize((size_t)nz); val.resize((size_t)nz);
    for (long long e = 0; e < nz; ++e) col[(size_t)e] = get_int(is, int_bytes);
    for (long long e = 0; e < nz; ++e) { const unsigned long long bits = get_be(is, 8); std::memcpy(&val[(size_t)e], &bits, 8); }
    return is.good() ? 0 : PETSC_ERR_ARG_CORRUPT;
}

/** SaveAndDestroy / SaveBlockInfo without the destroy (src/DMRGBlock.cpp:924-1012): the block stays in HBM */
inline PetscErrorCode Save(const Block::SpinBase& blk, const std::string& dir_in, int int_bytes = 4, int spin_type_key = 102) {
    std::string dir = dir_in;
    if (dir.empty() || dir.back() != '/') dir += '/';
    PetscErrorCode ierr = Makedir(dir); CHKERRQ(ierr);
    const PetscInt n = blk.NumStates();
    auto save_op = [&](int op, PetscInt isite, const std::string& name) -> PetscErrorCode {
        dmrgx_int nnz = 0;
        DMRGX_CALL(dmrgx_block_get_operator(blk.Handle(), op, isite, &nnz, NULL, NULL, NULL));
        std::vector<PetscInt> rp((size_t)n + 1), ci((size_t)std::max<dmrgx_int>(nnz, 1));
        std::vector<PetscScalar> vv((size_t)std::max<dmrgx_int>(nnz, 1));
        DMRGX_CALL(dmrgx_block_get_operator(blk.Handle(), op, isite, &nnz, rp.data(), ci.data(), vv.data()));
        return WriteMat(OpFilename(dir, name, (size_t)isite), n, rp, ci, vv, int_bytes);
    };
    for (PetscInt i = 0; i < blk.NumSites(); ++i) { ierr = save_op(DMRGX_OP_SZ, i, "Sz"); CHKERRQ(ierr); }
    for (PetscInt i = 0; i < blk.NumSites(); ++i) { ierr = save_op(DMRGX_OP_SP, i, "Sp"); CHKERRQ(ierr); }
    ierr = save_op(DMRGX_OP_H, 0, "H"); CHKERRQ(ierr);
    {
        std::ofstream f((dir + "BlockInfo.dat").c_str());
#define SaveInfo(KEY, VALUE) f << std::left << std::setfill(' ') << std::setw(30) << KEY << " " << VALUE << std::endl;
        SaveInfo("NumBytesPetscInt", int_bytes);
        SaveInfo("NumBytesPetscScalar", 8);
        SaveInfo("PetscUseComplex", 0);
        SaveInfo("SpinTypeKey", spin_type_key);
        SaveInfo("NumSites", blk.NumSites());
        SaveInfo("NumStates", blk.NumStates());
        SaveInfo("NumSectors", blk.Magnetization.NumSectors());
#undef SaveInfo
    }
    {
        std::ofstream f((dir + "QuantumNumbers.dat").c_str());
        for (PetscInt i = 0; i < blk.Magnetization.NumSectors(); ++i) f << blk.Magnetization.Sizes()[(size_t)i] << " " << blk.Magnetization.List()[(size_t)i] << std::endl;
    }
    return 0;
}

/** The spin type recorded in a block file against the -spin option (src/DMRGBlock.cpp:280-315): a mismatch is an error;
    without -spin the file's type is imposed as the option, so that sites added later have the spin of the loaded blocks. */
inline PetscErrorCode ApplySpinTypeKey(long long key) {
    const char* from_file = key == 102 ? "1/2" : (key == 101 ? "1" : nullptr); /* SpinOneHalf = 102, SpinOne = 101 (include/DMRGBlock.hpp:51-55) */
    if (!from_file) SETERRQ1(0, 1, "Input SpinTypeKey %lld not valid/implemented.", key);
    std::string spin; PetscBool set;
    PetscOptions::DB().GetString("-spin", spin, &set);
    if (set) {
        if (spin != "1/2" && spin != "1") SETERRQ1(0, 1, "Given -spin %s not valid/implemented.", spin.c_str());
        if (spin != from_file) SETERRQ2(0, 1, "Given spin types from file (%s) and command line (%s) do not match", from_file, spin.c_str());
    } else {
        PetscOptions::DB().kv["-spin"] = from_file;
    }
    return 0;
}
/** SpinTypeKey of a block directory (BlockInfo.dat), applied as above: -restart_dir calls this BEFORE the single site is created */
inline PetscErrorCode ApplySpinTypeOfBlockDir(const std::string& dir_in) {
    std::string dir = dir_in;
    if (dir.empty() || dir.back() != '/') dir += '/';
    std::ifstream info((dir + "BlockInfo.dat").c_str());
    if (!info) SETERRQ1(0, 1, "Error in reading %s", (dir + "BlockInfo.dat").c_str());
    for (std::string line; std::getline(info, line);) {
        std::istringstream iss(line); std::string k; long long v;
        if (iss >> k >> v && k == "SpinTypeKey") return ApplySpinTypeKey(v);
    }
    SETERRQ(0, 1, "SpinTypeKey not found.");
}

/** InitializeFromDisk (src/DMRGBlock.cpp:214-372) */
inline PetscErrorCode Load(Block::SpinBase& blk, const std::string& dir_in) {
    std::string dir = dir_in;
    if (dir.empty()) dir = "/";
    else if (dir.back() != '/') dir += '/';
    std::ifstream info((dir + "BlockInfo.dat").c_str());
    if (!info) SETERRQ1(0, 1, "Error in reading %s", (dir + "BlockInfo.dat").c_str());
    std::map<std::string, long long> m;
    for (std::string line; std::getline(info, line);) { std::istringstream iss(line); std::string k; long long v; if (iss >> k >> v) m[k] = v; }
    for (const char* k : {"NumBytesPetscInt", "NumBytesPetscScalar", "PetscUseComplex", "NumSites", "NumStates", "NumSectors"})
        if (!m.count(k)) SETERRQ1(0, 1, "BlockInfo.dat: %s not found.", k);
    if (!m.count("SpinTypeKey")) SETERRQ(0, 1, "SpinTypeKey not found.");
    { PetscErrorCode e = ApplySpinTypeKey(m["SpinTypeKey"]); CHKERRQ(e); }
    const int int_bytes = (int)m["NumBytesPetscInt"];
    if (int_bytes != 4 && int_bytes != 8) SETERRQ1(0, 1, "Incompatible NumBytesPetscInt. Got %lld.", m["NumBytesPetscInt"]);
    if (m["NumBytesPetscScalar"] != 8) SETERRQ1(0, 1, "Incompatible NumBytesPetscScalar. Expected 8. Got %lld.", m["NumBytesPetscScalar"]);
    if (m["PetscUseComplex"] != 0) SETERRQ1(0, 1, "Incompatible PetscUseComplex. Expected 0. Got %lld.", m["PetscUseComplex"]);
    const PetscInt nsec = m["NumSectors"];
    if (nsec <= 0) SETERRQ(0, 1, "NumSectors cannot be zero.");
    std::vector<PetscReal> qn; std::vector<PetscInt> sz;
    {
        std::ifstream f((dir + "QuantumNumbers.dat").c_str());
        if (!f) SETERRQ1(0, 1, "Error in reading %s", (dir + "QuantumNumbers.dat").c_str());
        for (std::string line; std::getline(f, line);) { std::istringstream iss(line); PetscInt s; PetscReal q; if (iss >> s >> q) { sz.push_back(s); qn.push_back(q); } }
        if ((PetscInt)qn.size() != nsec) SETERRQ2(0, 1, "Incorrect number of data points in QuantumNumbers.dat. Expected %lld. Got %lld.", LLD(nsec), LLD(qn.size()));
    }
    PetscErrorCode ierr = blk.Initialize(0, m["NumSites"], qn, sz); CHKERRQ(ierr);
    if (blk.NumStates() != m["NumStates"]) SETERRQ(0, 1, "NumStates of BlockInfo.dat does not match QuantumNumbers.dat");
    std::vector<PetscInt> rp, ci; std::vector<PetscScalar> vv;
    for (PetscInt i = 0; i < blk.NumSites(); ++i) {
        ierr = ReadMat(OpFilename(dir, "Sz", (size_t)i), blk.NumStates(), rp, ci, vv, int_bytes); CHKERRQ(ierr);
        ierr = blk.MatSetFromCSR(DMRGX_OP_SZ, i, rp.data(), ci.data(), vv.data()); CHKERRQ(ierr);
        ierr = ReadMat(OpFilename(dir, "Sp", (size_t)i), blk.NumStates(), rp, ci, vv, int_bytes); CHKERRQ(ierr);
        ierr = blk.MatSetFromCSR(DMRGX_OP_SP, i, rp.data(), ci.data(), vv.data()); CHKERRQ(ierr);
    }
    { /* H is optional on disk (:1073-1076) */
        std::ifstream probe(OpFilename(dir, "H", 0).c_str());
        if (probe) {
            ierr = ReadMat(OpFilename(dir, "H", 0), blk.NumStates(), rp, ci, vv, int_bytes); CHKERRQ(ierr);
            ierr = blk.MatSetFromCSR(DMRGX_OP_H, 0, rp.data(), ci.data(), vv.data()); CHKERRQ(ierr);
        }
    }
    return blk.CheckOperatorBlocks();
}

}  // namespace BlockIO

inline PetscErrorCode Block::SpinBase::InitializeFromDisk(const MPI_Comm&, const std::string& block_path) { return BlockIO::Load(*this, block_path); }
