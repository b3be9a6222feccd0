/*  PetscShim.hpp — the sliver of PETSc's vocabulary the reference's host code is written in, so that the
 *  classes below keep the reference's signatures (PetscInt, PetscErrorCode, CHKERRQ, the options database read with
 *  PetscOptionsGet*).  No PETSc, no MPI: one process drives one GPU context through the C ABI (include/dmrgx.h).
 */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

typedef long long PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef bool PetscBool;
#define PETSC_TRUE true
#define PETSC_FALSE false
#define PETSC_DEFAULT (-2)
#define PETSC_MAX_PATH_LEN 4096
typedef int MPI_Comm;
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 0
#define LLD(x) ((long long)(x))

/* petscerror.h (3.8) codes used on this path */
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_ARG_CORRUPT 64
#define PETSC_ERR_ARG_WRONGSTATE 73

#define CHKERRQ(ierr) do { if (ierr) { fprintf(stderr, "[dmrgx] error %d at %s:%d\n", (int)(ierr), __FILE__, __LINE__); return (ierr); } } while (0)
#define SETERRQ(comm, code, msg) do { fprintf(stderr, "[dmrgx] %s:%d: %s\n", __FILE__, __LINE__, (msg)); return (code); } while (0)
#define SETERRQ1(comm, code, fmt, a) do { fprintf(stderr, "[dmrgx] %s:%d: " fmt "\n", __FILE__, __LINE__, (a)); return (code); } while (0)
#define SETERRQ2(comm, code, fmt, a, b) do { fprintf(stderr, "[dmrgx] %s:%d: " fmt "\n", __FILE__, __LINE__, (a), (b)); return (code); } while (0)

/** The options database: `-key value` pairs from argv (and from `-options_file`), read ad hoc like PetscOptionsGet*. */
class PetscOptions {
public:
    static PetscOptions& DB() { static PetscOptions db; return db; }
    void Insert(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) {
            std::string k = argv[i];
            if (k.size() < 2 || k[0] != '-' || (k[1] >= '0' && k[1] <= '9')) continue;
            std::string v;
            if (i + 1 < argc) {
                std::string nx = argv[i + 1];
                const bool is_key = nx.size() >= 2 && nx[0] == '-' && !((nx[1] >= '0' && nx[1] <= '9') || nx[1] == '.');
                if (!is_key) { v = nx; ++i; }
            }
            kv[k] = v;
        }
    }
    void InsertString(const std::string& text) {
        std::vector<std::string> tok;
        std::istringstream iss(text);
        for (std::string t; iss >> t;) tok.push_back(t);
        std::vector<char*> av = {(char*)"x"};
        for (auto& t : tok) av.push_back((char*)t.c_str());
        Insert((int)av.size(), av.data());
    }
    bool Has(const std::string& k) const { return kv.count(k) > 0; }
    PetscErrorCode GetString(const char* key, std::string& out, PetscBool* set) const {
        auto f = kv.find(key);
        if (set) *set = (f != kv.end());
        if (f != kv.end()) out = f->second;
        return 0;
    }
    PetscErrorCode GetInt(const char* key, PetscInt* v, PetscBool* set) const {
        auto f = kv.find(key);
        if (set) *set = (f != kv.end());
        if (f != kv.end()) *v = atoll(f->second.c_str());
        return 0;
    }
    PetscErrorCode GetReal(const char* key, PetscReal* v, PetscBool* set) const {
        auto f = kv.find(key);
        if (set) *set = (f != kv.end());
        if (f != kv.end()) *v = atof(f->second.c_str());
        return 0;
    }
    PetscErrorCode GetBool(const char* key, PetscBool* v, PetscBool* set) const {
        auto f = kv.find(key);
        if (set) *set = (f != kv.end());
        if (f != kv.end()) {
            const std::string& s = f->second;
            *v = !(s == "0" || s == "false" || s == "no" || s == "FALSE" || s == "NO");
        }
        return 0;
    }
    PetscErrorCode GetIntArray(const char* key, std::vector<PetscInt>& out, PetscBool* set) const {
        auto f = kv.find(key);
        if (set) *set = (f != kv.end());
        if (f != kv.end()) {
            out.clear();
            std::string s = f->second;
            for (char& c : s) if (c == ',') c = ' ';
            std::istringstream iss(s);
            for (PetscInt x; iss >> x;) out.push_back(x);
        }
        return 0;
    }
    std::map<std::string, std::string> kv;
};

inline PetscErrorCode Makedir(const std::string& dir) {
    std::string cmd = "mkdir -p '" + dir + "'";
    return system(cmd.c_str()) ? 1 : 0;
}
