"""Pins the oracle's sector / index bookkeeping against the reference's own golden vectors.

tests/golden/testkron01.json is extracted (tests/golden/make_golden.py) from
/root/reference/tests/UnitTests_DMRGKron.cpp:39-252 (TestKron01) — the only golden rows the reference
holds for QuantumNumbers / KronBlocks_t / KronEye_Explicit index maps — and
tests/golden/opblocks.json from tests/UnitTests_DMRGBlock.cpp:76-131.
"""
import json
import os

import numpy as np
import pytest


def _block_from_fixture(O, spec):
    blk = O.Block.create(spec["nsites"], spec["qn"], spec["sizes"])
    n = blk.nstates
    for op, code in (("Sz", O.OP_SZ), ("Sp", O.OP_SP)):
        for site in range(spec["nsites"]):
            rows = [[] for _ in range(n)]
            for r in spec["rows"]:
                if r["op"] == op and r["site"] == site:
                    # SetRow stores value == column index (tests/UnitTests_Misc.cpp:15-18)
                    rows[r["row"]] = [(c, float(c)) for c in r["cols"]]
            blk.set_op_rows(code, site, rows)
    return blk


def test_testkron01_golden_rows(orc, golden_dir):
    O = orc
    fx = json.load(open(os.path.join(golden_dir, "testkron01.json")))
    L = _block_from_fixture(O, fx["blocks"]["Left"])
    R = _block_from_fixture(O, fx["blocks"]["Right"])
    out = O.kron_eye(L, R, [])
    assert out.nsites == 5 and out.nstates == 12
    assert out.check() == 0
    ops = {}
    for e in fx["expected"]:
        key = (e["op"], e["site"])
        if key not in ops:
            ops[key] = out.get_op(O.OP_SZ if e["op"] == "Sz" else O.OP_SP, e["site"])
        rowptr, col, val = ops[key]
        r = e["row"]
        got_c = col[rowptr[r]:rowptr[r + 1]].tolist()
        got_v = val[rowptr[r]:rowptr[r + 1]].tolist()
        assert got_c == e["cols"], (key, r, got_c, e["cols"])
        assert got_v == e["vals"], (key, r, got_v, e["vals"])
    assert len(ops) == 10  # all 10 output operators are pinned


def test_testkron01_sectors(orc, golden_dir):
    """Merged sector list of the 12-state block: L {+.5:2,-.5:1} x R {+1:1,0:2,-1:1}."""
    O = orc
    fx = json.load(open(os.path.join(golden_dir, "testkron01.json")))
    L = _block_from_fixture(O, fx["blocks"]["Left"])
    R = _block_from_fixture(O, fx["blocks"]["Right"])
    out = O.kron_eye(L, R, [])
    qn, sz = out.sectors()
    assert qn.tolist() == [1.5, 0.5, -0.5, -1.5]
    assert sz.tolist() == [2, 5, 4, 1]
    kb = O.KronBlocks(L, R, [])
    q, il, ir, size, off = kb.data()
    # IL-major enumeration, then stable sort by descending total QN (include/DMRGKron.hpp:147-158)
    assert list(zip(q.tolist(), il.tolist(), ir.tolist(), size.tolist())) == [
        (1.5, 0, 0, 2), (0.5, 0, 1, 4), (0.5, 1, 0, 1), (-0.5, 0, 2, 2), (-0.5, 1, 1, 2), (-1.5, 1, 2, 1)]
    assert off.tolist() == [0, 2, 6, 7, 9, 11, 12]
    assert kb.map(1, 1) == 4 and kb.map(2, 0) == -1 and kb.offsets_lr(5, 5) == -1


def test_check_operator_blocks_fixture(orc, golden_dir):
    O = orc
    fx = json.load(open(os.path.join(golden_dir, "opblocks.json")))
    blk = O.Block.create(2, fx["sectors"]["qn"], fx["sectors"]["sizes"])
    n = blk.nstates
    assert n == 8
    for name, rowsspec in fx["check_cases"].items():
        op, site = name[:2], int(name[3])
        rows = [[] for _ in range(n)]
        for r in rowsspec:
            rows[r["row"]] = [(c, float(c)) for c in r["cols"]]
        code = O.OP_SZ if op == "Sz" else O.OP_SP
        blk.set_op_rows(code, site, rows)
        shift = O.OpSz if op == "Sz" else O.OpSp
        assert blk.check_op(shift, code, site) == fx["check_expect"][name], name


def test_quantum_numbers_errors(orc):
    O = orc
    with pytest.raises(O.OracleError):
        O.Block.create(2, [0.5, 0.5], [1, 1])  # not strictly descending (src/QuantumNumbers.cpp:31-39)
    with pytest.raises(O.OracleError):
        O.Block.create(2, [], [])


def test_symeig_matches_numpy(orc):
    import ctypes as C
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 40):
        A = rng.standard_normal((n, n)); A = A + A.T
        w = np.zeros(n); V = np.zeros((n, n))
        assert orc.lib().orc_symeig(C.c_longlong(n), A.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                    V.ctypes.data_as(C.c_void_p)) == 0
        wn = np.linalg.eigvalsh(A)[::-1]
        assert np.allclose(w, wn, atol=1e-12)
        assert np.allclose(A @ V, V * w, atol=1e-11)


def test_oracle_reproduces_frozen_step(orc, golden_dir):
    """tests/golden/step_j1j2_4x4.json (tests/golden/make_step_golden.py) guards the oracle against drift: rebuilt from the stored
    input block, the oracle gives the same H·x, energy and truncation."""
    O = orc
    fx = json.load(open(os.path.join(golden_dir, "step_j1j2_4x4.json")))
    ham = fx["ham"]
    blk = O.Block.create(fx["nsites"], fx["qn"], fx["sizes"])
    for i in range(fx["nsites"]):
        for name, code in (("Sz", O.OP_SZ), ("Sp", O.OP_SP)):
            o = fx["ops"]["%s%d" % (name, i)]
            blk.set_op(code, i, np.array(o["rowptr"]), np.array(o["col"]), np.array(o["val"]))
    o = fx["ops"]["H"]
    blk.set_op(O.OP_H, 0, np.array(o["rowptr"]), np.array(o["col"]), np.array(o["val"]))
    T = lambda n: O.ham_terms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, ham["bcx"], ham["bcy"])
    enl = O.kron_eye(blk, O.Block.single_site(), T(8))
    kb = O.KronBlocks(enl, enl, [0.0])
    sh = O.Shell(kb, T(16))
    y = sh.apply(np.array(fx["x"]))
    assert np.abs(y - np.array(fx["y"])).max() <= 1e-14 * np.abs(y).max()
    e0, psi, _, _ = sh.eigs(tol=1e-13)
    assert abs(e0 - fx["e0"]) <= 1e-11 * abs(e0)
    tL = O.Truncation(kb, np.array(fx["psi"]), fx["mstates"], True)
    assert tL.sectors()[1].tolist() == fx["trunc"]["L"]["sizes"] and abs(tL.trunc_err - fx["trunc"]["L"]["err"]) < 1e-13
