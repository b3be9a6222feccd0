"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liboracle.so`` (the CPU restatement of the reference's hot path,
``oracle/dmrg_oracle.hpp``).  Imported only by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs — never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OpSm, OpSz, OpSp, OpEye = -1, 0, 1, 2
OP_SZ, OP_SP, OP_H = 0, 1, 3  # operator selectors of the block accessors

LL = C.c_longlong
PLL = C.POINTER(C.c_longlong)
PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_block_single_site", "orc_block_create", "orc_kron_eye", "orc_kron_create", "orc_shell_create", "orc_shell_create_rows",
                     "orc_shell_create_single", "orc_truncate", "orc_rotate", "orc_dmrg_create", "orc_dmrg_block"):
            getattr(L, name).restype = C.c_void_p
        for name in ("orc_block_op_nnz", "orc_ham_terms", "orc_kron_size", "orc_kron_num_states", "orc_kron_map",
                     "orc_kron_offsets_lr", "orc_shell_nterms", "orc_shell_lrows", "orc_dmrg_nsteps", "orc_dmrg_step_nsectors"):
            getattr(L, name).restype = LL
        L.orc_shell_fmas.restype = C.c_double
        L.orc_eigs.restype = C.c_double
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _l(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__("oracle error %d: %s" % (code, lib().orc_last_error().decode()))
        self.code = code


def terms_arrays(terms):
    """terms: list of (a, Iop, Isite, Jop, Jsite)"""
    a = _d([t[0] for t in terms])
    iop = _i([t[1] for t in terms])
    isite = _l([t[2] for t in terms])
    jop = _i([t[3] for t in terms])
    jsite = _l([t[4] for t in terms])
    return a, iop, isite, jop, jsite


def ham_terms(Lx, Ly, J1, Jz1, J2, Jz2, nsites, bcx=0, bcy=1):
    n = 8 * max(nsites, Lx * Ly) * 3 + 16
    a = np.zeros(n); iop = np.zeros(n, np.int32); isite = np.zeros(n, np.int64)
    jop = np.zeros(n, np.int32); jsite = np.zeros(n, np.int64)
    k = lib().orc_ham_terms(LL(Lx), LL(Ly), C.c_double(J1), C.c_double(Jz1), C.c_double(J2), C.c_double(Jz2), bcx, bcy, LL(nsites),
                            LL(n), _p(a), _p(iop), _p(isite), _p(jop), _p(jsite))
    assert k <= n
    return [(float(a[i]), int(iop[i]), int(isite[i]), int(jop[i]), int(jsite[i])) for i in range(k)]


class Block:
    def __init__(self, handle, owned=True):
        self.h = C.c_void_p(handle)
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            lib().orc_block_free(self.h)
            self.h = None

    @staticmethod
    def single_site(spin_twice=1):
        return Block(lib().orc_block_single_site(spin_twice))

    @staticmethod
    def create(nsites, qn, sizes):
        qn = _d(qn); sizes = _l(sizes)
        err = C.c_int(0)
        h = lib().orc_block_create(LL(nsites), LL(len(qn)), _p(qn), _p(sizes), C.byref(err))
        if err.value:
            raise OracleError(err.value)
        return Block(h)

    def info(self):
        a, b, c = LL(), LL(), LL()
        lib().orc_block_info(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    @property
    def nsites(self):
        return self.info()[0]

    @property
    def nstates(self):
        return self.info()[1]

    def sectors(self):
        ns = self.info()[2]
        qn = np.zeros(ns); sz = np.zeros(ns, np.int64)
        lib().orc_block_sectors(self.h, _p(qn), _p(sz))
        return qn, sz

    def set_op(self, optype, isite, rowptr, col, val):
        rowptr = _l(rowptr); col = _l(col); val = _d(val)
        e = lib().orc_block_set_op(self.h, optype, LL(isite), _p(rowptr), _p(col), _p(val))
        if e:
            raise OracleError(e)

    def set_op_rows(self, optype, isite, rows):
        """rows: list (len nstates) of lists of (col, val)"""
        rowptr = [0]; col = []; val = []
        for r in rows:
            for c, v in sorted(r):
                col.append(c); val.append(v)
            rowptr.append(len(col))
        self.set_op(optype, isite, rowptr, col, val)

    def get_op(self, optype, isite=0):
        n = self.nstates
        nnz = lib().orc_block_op_nnz(self.h, optype, LL(isite))
        if nnz < 0:
            raise IndexError("no such operator")
        rowptr = np.zeros(n + 1, np.int64); col = np.zeros(max(nnz, 1), np.int64); val = np.zeros(max(nnz, 1))
        lib().orc_block_get_op(self.h, optype, LL(isite), _p(rowptr), _p(col), _p(val))
        return rowptr, col[:nnz], val[:nnz]

    def get_op_dense(self, optype, isite=0):
        rowptr, col, val = self.get_op(optype, isite)
        n = self.nstates
        D = np.zeros((n, n))
        for r in range(n):
            for k in range(rowptr[r], rowptr[r + 1]):
                D[r, col[k]] += val[k]
        return D

    def check(self):
        return lib().orc_block_check(self.h)

    def check_op(self, shift, optype, isite):
        return lib().orc_block_check_op(self.h, shift, optype, LL(isite))


def kron_eye(L, R, terms):
    a, iop, isite, jop, jsite = terms_arrays(terms)
    err = C.c_int(0)
    h = lib().orc_kron_eye(L.h, R.h, len(terms), _p(a), _p(iop), _p(isite), _p(jop), _p(jsite), C.byref(err))
    if err.value:
        raise OracleError(err.value)
    return Block(h)


class KronBlocks:
    def __init__(self, L, R, qn_sectors):
        self.L, self.R = L, R
        qn = _d(qn_sectors)
        err = C.c_int(0)
        self.h = C.c_void_p(lib().orc_kron_create(L.h, R.h, len(qn), _p(qn), C.byref(err)))
        if err.value:
            raise OracleError(err.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kron_free(self.h)
            self.h = None

    def size(self):
        return lib().orc_kron_size(self.h)

    def num_states(self):
        return lib().orc_kron_num_states(self.h)

    def data(self):
        n = self.size()
        qn = np.zeros(n); il = np.zeros(n, np.int64); ir = np.zeros(n, np.int64); sz = np.zeros(n, np.int64)
        off = np.zeros(n + 1, np.int64)
        lib().orc_kron_data(self.h, _p(qn), _p(il), _p(ir), _p(sz), _p(off))
        return qn, il, ir, sz, off

    def map(self, l, r):
        return lib().orc_kron_map(self.h, LL(l), LL(r))

    def offsets_lr(self, l, r):
        return lib().orc_kron_offsets_lr(self.h, LL(l), LL(r))


class Shell:
    def __init__(self, kb, terms=None, single=None, rows=None):
        """rows=(r0, r1): build only that row range (what one MPI rank of the reference owns); apply() then returns
        the lrows values of those rows."""
        self.kb = kb
        err = C.c_int(0)
        if rows is not None:
            a, iop, isite, jop, jsite = terms_arrays(terms)
            self.h = C.c_void_p(lib().orc_shell_create_rows(kb.h, len(terms), _p(a), _p(iop), _p(isite), _p(jop), _p(jsite),
                                                            LL(rows[0]), LL(rows[1]), C.byref(err)))
        elif single is not None:
            opl, il, opr, ir = single
            self.h = C.c_void_p(lib().orc_shell_create_single(kb.h, opl, LL(il), opr, LL(ir), C.byref(err)))
        else:
            a, iop, isite, jop, jsite = terms_arrays(terms)
            self.h = C.c_void_p(lib().orc_shell_create(kb.h, len(terms), _p(a), _p(iop), _p(isite), _p(jop), _p(jsite), C.byref(err)))
        if err.value:
            raise OracleError(err.value)
        self.n = kb.num_states()
        self.lrows = lib().orc_shell_lrows(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_shell_free(self.h)
            self.h = None

    def nterms(self):
        return lib().orc_shell_nterms(self.h)

    def fmas(self):
        return lib().orc_shell_fmas(self.h)

    def apply(self, x, nthreads=1):
        x = _d(x); y = np.zeros(self.lrows)
        lib().orc_shell_apply_rows(self.h, _p(x), _p(y), LL(0), LL(self.lrows), int(nthreads))
        return y

    def apply_rows(self, x, r0, r1, nthreads=1, y=None):
        x = _d(x)
        if y is None:
            y = np.zeros(self.n)
        lib().orc_shell_apply_rows(self.h, _p(x), _p(y), LL(r0), LL(r1), nthreads)
        return y

    def eigs(self, tol=1e-12, ncv=16, max_it=2000):
        psi = np.zeros(self.n); nmv = LL(); res = C.c_double()
        e = lib().orc_eigs(self.h, C.c_double(tol), LL(ncv), LL(max_it), _p(psi), C.byref(nmv), C.byref(res))
        return e, psi, nmv.value, res.value


class Truncation:
    def __init__(self, kb, psi, mstates, left):
        psi = _d(psi)
        err = C.c_int(0)
        self.h = C.c_void_p(lib().orc_truncate(kb.h, _p(psi), LL(mstates), int(left), C.byref(err)))
        if err.value:
            raise OracleError(err.value)
        m, n, ns, te, tie, nspec = LL(), LL(), LL(), C.c_double(), C.c_int(), LL()
        lib().orc_bt_info(self.h, C.byref(m), C.byref(n), C.byref(ns), C.byref(te), C.byref(tie), C.byref(nspec))
        self.m, self.nstates, self.nsectors, self.trunc_err, self.tie, self.nspec = m.value, n.value, ns.value, te.value, bool(tie.value), nspec.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bt_free(self.h)
            self.h = None

    def sectors(self):
        qn = np.zeros(self.nsectors); sz = np.zeros(self.nsectors, np.int64)
        lib().orc_bt_sectors(self.h, _p(qn), _p(sz))
        return qn, sz

    def spectrum(self):
        ev = np.zeros(self.nspec); bi = np.zeros(self.nspec, np.int64)
        lib().orc_bt_spectrum(self.h, _p(ev), _p(bi))
        return ev, bi

    def rotmat(self):
        out = np.zeros((self.m, self.nstates))
        lib().orc_bt_rotmat_dense(self.h, _p(out))
        return out


def rotate(blk_enl, bt):
    err = C.c_int(0)
    h = lib().orc_rotate(blk_enl.h, bt.h, C.byref(err))
    if err.value:
        raise OracleError(err.value)
    return Block(h)


STEP_INT_FIELDS = ["GlobIdx", "LoopType", "LoopIdx", "StepIdx", "NSites_Sys", "NSites_Env", "NSites_SysEnl", "NSites_EnvEnl",
                   "NStates_Sys", "NStates_Env", "NStates_SysEnl", "NStates_EnvEnl", "NStates_SysRot", "NStates_EnvRot", "NumStates_H"]


class DMRG:
    def __init__(self, Lx, Ly, J1=1.0, Jz1=0.0, J2=1.0, Jz2=0.0, bcx=0, bcy=1, heisenberg=None, spin_twice=1, qn_sector=0.0,
                 eps_tol=1e-12, eps_ncv=16, eps_max_it=2000):
        if heisenberg is not None:  # include/Hamiltonians.hpp:102-108
            Jz1, J1, J2, Jz2 = heisenberg, 0.5, 0.0, 0.0
        self.params = dict(Lx=Lx, Ly=Ly, J1=J1, Jz1=Jz1, J2=J2, Jz2=Jz2, bcx=bcx, bcy=bcy)
        self.h = C.c_void_p(lib().orc_dmrg_create(LL(Lx), LL(Ly), C.c_double(J1), C.c_double(Jz1), C.c_double(J2), C.c_double(Jz2),
                                                  bcx, bcy, spin_twice, C.c_double(qn_sector), C.c_double(eps_tol), LL(eps_ncv),
                                                  LL(eps_max_it)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_dmrg_free(self.h)
            self.h = None

    def warmup(self, mwarmup):
        e = lib().orc_dmrg_warmup(self.h, LL(mwarmup))
        if e:
            raise OracleError(e)

    def sweep(self, mstates):
        e = lib().orc_dmrg_sweep(self.h, LL(mstates))
        if e:
            raise OracleError(e)

    def steps(self):
        out = []
        for i in range(lib().orc_dmrg_nsteps(self.h)):
            ints = np.zeros(15, np.int64); reals = np.zeros(3); flags = np.zeros(3, np.int64)
            lib().orc_dmrg_step(self.h, LL(i), _p(ints), _p(reals), _p(flags))
            d = {k: int(v) for k, v in zip(STEP_INT_FIELDS, ints)}
            d.update(TruncErr_Sys=float(reals[0]), TruncErr_Env=float(reals[1]), GSEnergy=float(reals[2]), nmatvec=int(flags[0]),
                     tie_L=bool(flags[1]), tie_R=bool(flags[2]))
            for side, key in ((1, "L"), (0, "R")):
                ns = lib().orc_dmrg_step_nsectors(self.h, LL(i), side)
                qn = np.zeros(ns); sz = np.zeros(ns, np.int64)
                lib().orc_dmrg_step_sectors(self.h, LL(i), side, _p(qn), _p(sz))
                d["qn_list_" + key] = qn.tolist(); d["qn_size_" + key] = sz.tolist()
            out.append(d)
        return out

    def block(self, i):
        h = lib().orc_dmrg_block(self.h, LL(i))
        return Block(h, owned=False) if h else None
