/*  Hamiltonians::J1J2XXZModel_SquareLattice — host mirror of include/Hamiltonians.hpp:77-288.
 *  The term list itself comes from the library (dmrgx_ham_terms <-> src/Hamiltonians.cpp:73-122). */
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmrgx.h"
#include "PetscShim.hpp"

typedef enum { OpSm = -1, OpSz = 0, OpSp = +1, OpEye = +2 } Op_t; /* include/DMRGBlock.hpp:21-27 */

namespace Hamiltonians {

struct Term { PetscScalar a; Op_t Iop; PetscInt Isite; Op_t Jop; PetscInt Jsite; };
typedef enum { OpenBC = 0, PeriodicBC = 1 } BC_t;

class J1J2XXZModel_SquareLattice {
public:
    /** include/Hamiltonians.hpp:89-118 */
    PetscErrorCode SetFromOptions() {
        PetscOptions& o = PetscOptions::DB();
        o.GetReal("-J1", &_J1, NULL); o.GetReal("-J2", &_J2, NULL); o.GetReal("-Jz1", &_Jz1, NULL); o.GetReal("-Jz2", &_Jz2, NULL);
        o.GetInt("-Lx", &_Lx, NULL); o.GetInt("-Ly", &_Ly, NULL);
        o.GetReal("-heisenberg", &_Jz1, &heisenberg);
        if (heisenberg) { _J1 = 0.50; _J2 = 0.0; _Jz2 = 0.0; }
        PetscBool BCopen = PETSC_FALSE, BCperiodic = PETSC_FALSE;
        o.GetBool("-BCopen", &BCopen, NULL);
        if (o.Has("-BCopen") && o.kv["-BCopen"].empty()) BCopen = PETSC_TRUE;
        if (BCopen) { _BCx = OpenBC; _BCy = OpenBC; }
        o.GetBool("-BCperiodic", &BCperiodic, NULL);
        if (o.Has("-BCperiodic") && o.kv["-BCperiodic"].empty()) BCperiodic = PETSC_TRUE;
        if (BCperiodic) { _BCx = PeriodicBC; _BCy = PeriodicBC; }
        return 0;
    }
    /** include/Hamiltonians.hpp:122-155 */
    PetscErrorCode SaveAsOptions(const std::string& filename) {
        FILE* fp = fopen(filename.c_str(), "w");
        if (!fp) return 1;
        for (const char* key : {"-J1", "-J2", "-Jz1", "-Jz2", "-Lx", "-Ly", "-heisenberg", "-BCopen", "-BCperiodic"}) {
            std::string v; PetscBool set;
            PetscOptions::DB().GetString(key, v, &set);
            if (set) fprintf(fp, "%s %s\n", key, v.empty() ? "yes" : v.c_str());
        }
        fclose(fp);
        return 0;
    }
    PetscInt NumSites() const { return _Lx * _Ly; }
    PetscInt NumEnvSites() const { return _Ly; }
    PetscInt Lx() const { return _Lx; }
    PetscInt Ly() const { return _Ly; }
    /** src/Hamiltonians.cpp:73-122 */
    std::vector<Term> H(const PetscInt& nsites) {
        const PetscInt cap = 16 * NumSites() * 3 + 16;
        std::vector<double> a(cap); std::vector<int> iop(cap), jop(cap); std::vector<dmrgx_int> is(cap), js(cap);
        const dmrgx_int n = dmrgx_ham_terms(_Lx, _Ly, _J1, _Jz1, _J2, _Jz2, (int)_BCx, (int)_BCy, nsites == PETSC_DEFAULT ? -1 : nsites, cap,
                                            a.data(), iop.data(), is.data(), jop.data(), js.data());
        std::vector<Term> t;
        for (dmrgx_int i = 0; i < n; ++i) t.push_back({a[i], (Op_t)iop[i], is[i], (Op_t)jop[i], js[i]});
        return t;
    }
    PetscInt To1D(const PetscInt ix, const PetscInt jy) const { return (ix % 2 == 0) ? ix * _Ly + jy : (ix + 1) * _Ly - (jy + 1); }
    PetscErrorCode To2D(const PetscInt idx, PetscInt& ix, PetscInt& jy) const {
        ix = idx / _Ly; jy = (ix % 2 == 0) ? idx % _Ly : _Ly - 1 - idx % _Ly; return 0;
    }
    /** src/Hamiltonians.cpp:124-147 — ordered 1-D index pairs of all nearest-neighbour bonds (up, then right, per site) */
    std::vector<std::vector<PetscInt>> NeighborPairs(const PetscInt d = 1) const {
        if (d != 1) throw std::runtime_error("Only d=1 supported.");
        std::vector<std::vector<PetscInt>> nnp;
        const PetscInt ns = _Lx * _Ly;
        for (PetscInt is = 0; is < ns; ++is) {
            PetscInt ix, jy;
            To2D(is, ix, jy);
            std::vector<PetscInt> nn;
            if (jy < _Ly - 1 || _BCy == PeriodicBC) { /* src/Hamiltonians.cpp:26-47 */
                const PetscInt ju = (jy + 1) % _Ly;
                if (ju != jy && To1D(ix, ju) < ns) nn.push_back(To1D(ix, ju));
            }
            if (ix < _Lx - 1 || _BCx == PeriodicBC) {
                const PetscInt ir = (ix + 1) % _Lx;
                if (ir != ix && To1D(ir, jy) < ns) nn.push_back(To1D(ir, jy));
            }
            for (const PetscInt in : nn) nnp.push_back({std::min(in, is), std::max(in, is)});
        }
        return nnp;
    }
    void PrintOut() const {
        printf("HAMILTONIAN: %s\n", heisenberg ? "HeisenbergModel_SquareLattice" : "J1J2XXZModel_SquareLattice");
        printf("  Lx  : %lld\n  Ly  : %lld\n  J1  : %g\n  Jz1 : %g\n  J2  : %g\n  Jz2 : %g\n  BCx : %s\n  BCy : %s\n", LLD(_Lx), LLD(_Ly), _J1, _Jz1,
               _J2, _Jz2, _BCx ? "Periodic" : "Open", _BCy ? "Periodic" : "Open");
    }
    /** include/Hamiltonians.hpp:190-215 — the "Hamiltonian" object of DMRGRun.json */
    void SaveOut(FILE* fp) const {
        fprintf(fp, "  \"Hamiltonian\": {\n");
        fprintf(fp, "    \"label\":\"%s\",\n", heisenberg ? "HeisenbergModel_SquareLattice" : "J1J2XXZModel_SquareLattice");
        fprintf(fp, "    \"parameters\": {\n");
        fprintf(fp, "      \"Lx\"  : %lld,\n", LLD(_Lx));
        fprintf(fp, "      \"Ly\"  : %lld,\n", LLD(_Ly));
        fprintf(fp, "      \"J1\"  : %g,\n", _J1);
        fprintf(fp, "      \"Jz1\" : %g,\n", _Jz1);
        fprintf(fp, "      \"J2\"  : %g,\n", _J2);
        fprintf(fp, "      \"Jz2\" : %g,\n", _Jz2);
        fprintf(fp, "      \"BCx\" : \"%s\",\n", _BCx ? "Periodic" : "Open");
        fprintf(fp, "      \"BCy\" : \"%s\"\n", _BCy ? "Periodic" : "Open");
        fprintf(fp, "    }\n");
        fprintf(fp, "  }");
        fflush(fp);
    }

private:
    PetscBool heisenberg = PETSC_FALSE;
    PetscScalar _Jz1 = 0.0, _J1 = 1.0, _Jz2 = 0.0, _J2 = 1.0; /* defaults: include/Hamiltonians.hpp:240-252 */
    PetscInt _Lx = 4, _Ly = 4;
    BC_t _BCx = OpenBC, _BCy = PeriodicBC; /* cylinder by default */
};

}  // namespace Hamiltonians
