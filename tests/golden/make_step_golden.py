"""Generates tests/golden/step_j1j2_4x4.json with the ORACLE (the reference itself cannot run here: no PETSc/SLEPc).

One SingleDMRGStep of the 4x4 J1-J2 cylinder (-J1 .5 -Jz1 1 -J2 .25 -Jz2 .5) frozen as numbers: the input block (sector list +
CSR of every operator, what InitializeFromDisk would hand over), a seeded vector x with y = H x, the ground-state energy, and the
truncation (kept-state counts per sector, truncation error, sorted spectrum) at a cut without ties.  The fixture guards the
oracle against drift and gives the CUDA path a target that does not depend on running the oracle at test time.
Run from the repo root:  python tests/golden/make_step_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HAM = dict(Lx=4, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5, bcx=0, bcy=1)


def terms(n):
    return O.ham_terms(HAM["Lx"], HAM["Ly"], HAM["J1"], HAM["Jz1"], HAM["J2"], HAM["Jz2"], n, HAM["bcx"], HAM["bcy"])


def main():
    d = O.DMRG(HAM["Lx"], HAM["Ly"], HAM["J1"], HAM["Jz1"], HAM["J2"], HAM["Jz2"], HAM["bcx"], HAM["bcy"], eps_tol=1e-13)
    d.warmup(20)
    blk = d.block(6)
    qn, sz = blk.sectors()
    ops = {}
    for i in range(blk.nsites):
        for name, code in (("Sz", O.OP_SZ), ("Sp", O.OP_SP)):
            rp, ci, vv = blk.get_op(code, i)
            ops["%s%d" % (name, i)] = dict(rowptr=rp.tolist(), col=ci.tolist(), val=[float(v) for v in vv])
    rp, ci, vv = blk.get_op(O.OP_H, 0)
    ops["H"] = dict(rowptr=rp.tolist(), col=ci.tolist(), val=[float(v) for v in vv])
    enl = O.kron_eye(blk, O.Block.single_site(), terms(8))
    kb = O.KronBlocks(enl, enl, [0.0])
    sh = O.Shell(kb, terms(16))
    rng = np.random.default_rng(20261018)
    x = rng.standard_normal(sh.n)
    y = sh.apply(x)
    e0, psi, _, _ = sh.eigs(tol=1e-13)
    for m in range(28, 40):
        tL = O.Truncation(kb, psi, m, True); tR = O.Truncation(kb, psi, m, False)
        if not (tL.tie or tR.tie):
            break
    assert not (tL.tie or tR.tie)
    q, il, ir, size, off = kb.data()
    out = dict(source="oracle/dmrg_oracle.hpp (restatement of include/DMRGBlockContainer.hpp:1304-1653 for one step)", ham=HAM, nsites=blk.nsites,
               qn=qn.tolist(), sizes=sz.tolist(), ops=ops,
               enlarged=dict(qn=enl.sectors()[0].tolist(), sizes=enl.sectors()[1].tolist()),
               kron=dict(qn=q.tolist(), il=il.tolist(), ir=ir.tolist(), size=size.tolist(), off=off.tolist()),
               x=[float(v) for v in x], y=[float(v) for v in y], e0=float(e0), psi=[float(v) for v in psi], mstates=m,
               trunc=dict(L=dict(qn=tL.sectors()[0].tolist(), sizes=tL.sectors()[1].tolist(), err=float(tL.trunc_err),
                                 spectrum=sorted((float(v) for v in tL.spectrum()[0]), reverse=True)),
                          R=dict(qn=tR.sectors()[0].tolist(), sizes=tR.sectors()[1].tolist(), err=float(tR.trunc_err),
                                 spectrum=sorted((float(v) for v in tR.spectrum()[0]), reverse=True))))
    path = os.path.join(HERE, "step_j1j2_4x4.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes; D =", sh.n, "E0 =", e0, "m =", m)


if __name__ == "__main__":
    main()
