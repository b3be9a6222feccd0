"""Shared parity checks: the product (through its C ABI) against the CPU oracle on the same inputs.

Used twice: by the GPU tests (tests/test_gpu_parity.py, real kernels) and by the CPU plan-check tests
(tests/test_plancheck_cpu.py, host planning logic interpreted by tests/plancheck/dev_host.cpp).
"""
import numpy as np

MATVEC_RTOL = 1e-13   # H·psi, relative to max|y|
ENERGY_RTOL = 1e-10   # north_star: energies and truncation errors within 1e-10 relative


def upload_block(P, ctx, oblk, O):
    """oracle Block -> product Block through dmrgx_block_set_operator (CSR, global indices)."""
    nsites, nstates, _ = oblk.info()
    qn, sz = oblk.sectors()
    b = P.Block.Initialize(ctx, nsites, qn, sz)
    for i in range(nsites):
        b.set_operator(P.OpSz, i, *oblk.get_op(O.OP_SZ, i))
        b.set_operator(P.OpSp, i, *oblk.get_op(O.OP_SP, i))
    b.set_operator(P.OpH, 0, *oblk.get_op(O.OP_H, 0))
    return b


def dense_op(blk, op, i=0):
    return blk.get_operator_dense(op, i)


def assert_blocks_equal(P, O, pblk, oblk, tol=1e-13, what=""):
    assert pblk.NumSites() == oblk.nsites and pblk.NumStates() == oblk.nstates, what
    pq, ps = pblk.sectors(); oq, os_ = oblk.sectors()
    assert pq.tolist() == oq.tolist() and ps.tolist() == os_.tolist(), what
    for i in range(oblk.nsites):
        for pop, oop in ((P.OpSz, O.OP_SZ), (P.OpSp, O.OP_SP)):
            a = dense_op(pblk, pop, i); b = oblk.get_op_dense(oop, i)
            assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (what, pop, i, np.abs(a - b).max())
    a = dense_op(pblk, P.OpH); b = oblk.get_op_dense(O.OP_H)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (what, "H", np.abs(a - b).max())


def check_kron_bookkeeping(P, O, pk, ok):
    """bit-exact: pair list, sizes, offsets, (IL,IR)->idx map"""
    pq, pil, pir, psz, poff = pk.data()
    oq, oil, oir, osz, ooff = ok.data()
    assert pq.tolist() == oq.tolist()
    assert pil.tolist() == oil.tolist() and pir.tolist() == oir.tolist()
    assert psz.tolist() == osz.tolist() and poff.tolist() == ooff.tolist()
    assert pk.NumStates() == ok.num_states() and pk.size() == ok.size()
    nl = int(pil.max()) + 2 if len(pil) else 1
    nr = int(pir.max()) + 2 if len(pir) else 1
    for l in range(-1, nl):
        for r in range(-1, nr):
            assert pk.Map(l, r) == ok.map(l, r)
            assert pk.Offsets(l, r) == ok.offsets_lr(l, r)


def check_matvec(P, O, ctx, pshell, oshell, rng, nvec=2):
    n = pshell.n
    assert n == oshell.n
    for _ in range(nvec):
        x = rng.standard_normal(n)
        y_ref = oshell.apply(x)
        y = pshell.MatMult_host(x)
        scale = max(np.abs(y_ref).max(), 1e-300)
        assert np.abs(y - y_ref).max() <= MATVEC_RTOL * scale * max(1, oshell.nterms()), np.abs(y - y_ref).max() / scale
        dx = ctx.vec(n, x); dy = ctx.vec(n)
        pshell.MatMult(dx, dy)
        assert np.array_equal(dy.get(), y)  # device-pointer and host-buffer entry points agree bit for bit


def check_truncation(P, O, pbt, obt, blk_sizes):
    """kept-state counts bit-exact; truncation error within 1e-10; rotation spans the same subspace"""
    pq, ps = pbt.sectors(); oq, os_ = obt.sectors()
    assert pq.tolist() == oq.tolist(), (pq, oq)
    assert ps.tolist() == os_.tolist(), (ps, os_)
    assert pbt.m == obt.m and pbt.nstates == obt.nstates
    assert abs(pbt.TruncErr - obt.trunc_err) <= ENERGY_RTOL * max(abs(obt.trunc_err), 1e-4), (pbt.TruncErr, obt.trunc_err)
    pe, pb = pbt.spectrum(); oe, ob = obt.spectrum()
    assert pb.tolist() == ob.tolist()
    assert np.abs(pe - oe).max() <= 1e-12
    Up = pbt.RotMatT(); Uo = obt.rotmat()
    # rows orthonormal
    assert np.abs(Up @ Up.T - np.eye(pbt.m)).max() < 1e-10
    # same projector when the cut is not inside a degenerate multiplet
    if not obt.tie:
        assert np.abs(Up.T @ Up - Uo.T @ Uo).max() < 1e-7


def lr_terms(O, ham, nsites):
    return O.ham_terms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], nsites, ham.get("bcx", 0), ham.get("bcy", 1))


def run_step_parity(P, O, ctx, ham, nsys, nenv, m_prep, m_keep, rng, tol=1e-12):
    """One SingleDMRGStep (include/DMRGBlockContainer.hpp:1304-1653) on both sides, from blocks the ORACLE prepared
    (warm-up with m_prep states): enlarge -> KronBlocks -> shell H -> solve -> truncate -> rotate."""
    d = O.DMRG(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], ham.get("bcx", 0), ham.get("bcy", 1), eps_tol=tol)
    d.warmup(m_prep)
    osys, oenv = d.block(nsys - 1), d.block(nenv - 1)
    assert osys is not None and oenv is not None
    osite = O.Block.single_site()
    # --- oracle side ---
    oL = O.kron_eye(osys, osite, lr_terms(O, ham, nsys + 1))
    oR = O.kron_eye(oenv, osite, lr_terms(O, ham, nenv + 1))
    okb = O.KronBlocks(oL, oR, [0.0])
    terms = lr_terms(O, ham, nsys + nenv + 2)
    osh = O.Shell(okb, terms)
    # --- product side: same input blocks, enlargement done by the product ---
    psys = upload_block(P, ctx, osys, O)
    penv = upload_block(P, ctx, oenv, O)
    psite = P.Block.SingleSite(ctx)
    pL = P.KronEye_Explicit(psys, psite, lr_terms(O, ham, nsys + 1))
    pR = P.KronEye_Explicit(penv, psite, lr_terms(O, ham, nenv + 1))
    assert_blocks_equal(P, O, pL, oL, what="enlarged sys")
    assert_blocks_equal(P, O, pR, oR, what="enlarged env")
    pkb = P.KronBlocks(pL, pR, [0.0])
    check_kron_bookkeeping(P, O, pkb, okb)
    psh = pkb.KronSumConstruct(terms)
    check_matvec(P, O, ctx, psh, osh, rng)
    # --- ground state ---
    e_ref, psi_ref, _, _ = osh.eigs(tol=tol)
    e, psi, st = psh.EPSSolve(tol=tol)
    assert st["converged"]
    assert abs(e - e_ref) <= ENERGY_RTOL * abs(e_ref), (e, e_ref)
    ph = psi.get()
    assert abs(abs(ph @ psi_ref) - 1.0) < 1e-8
    # --- truncation: feed BOTH sides the same vector so kept-state counts are comparable bit for bit ---
    psi_dev = ctx.vec(len(psi_ref), psi_ref)
    # Sz -> -Sz symmetry makes rho eigenvalues of the +q and -q sectors exactly degenerate, so a cut that falls
    # inside such a pair has no well-defined kept-state counts (the reference's answer depends on LAPACK round-off,
    # SURVEY.md §7 "Degenerate truncation cut"): move the cut to the next m at which the oracle reports no tie.
    for m_keep in range(m_keep, m_keep + 8):
        obtL = O.Truncation(okb, psi_ref, m_keep, True)
        obtR = O.Truncation(okb, psi_ref, m_keep, False)
        if not (obtL.tie or obtR.tie):
            break
    assert not (obtL.tie or obtR.tie)
    pbtL, pbtR = P.GetTruncation(pkb, psi_dev, m_keep)
    check_truncation(P, O, pbtL, obtL, None)
    check_truncation(P, O, pbtR, obtR, None)
    # --- rotation: compare gauge-invariant quantities ---
    pnew = P.RotateOperators(pL, pbtL)
    onew = O.rotate(oL, obtL)
    assert pnew.NumStates() == onew.nstates
    Up, Uo = pbtL.RotMatT(), obtL.rotmat()
    Hp = dense_op(pnew, P.OpH); Ho = onew.get_op_dense(O.OP_H)
    HL = oL.get_op_dense(O.OP_H)
    assert np.abs(Hp - Up @ HL @ Up.T).max() < 1e-11 * max(1.0, np.abs(HL).max())
    if not obtL.tie:
        assert np.abs(np.linalg.eigvalsh(Hp) - np.linalg.eigvalsh(Ho)).max() < 1e-9
    for i in range(onew.nsites):
        Sp_enl = oL.get_op_dense(O.OP_SP, i)
        assert np.abs(dense_op(pnew, P.OpSp, i) - Up @ Sp_enl @ Up.T).max() < 1e-11
        Sz_enl = oL.get_op_dense(O.OP_SZ, i)
        assert np.abs(dense_op(pnew, P.OpSz, i) - Up @ Sz_enl @ Up.T).max() < 1e-11
    assert pnew.CheckOperatorBlocks() == 0
    return e, e_ref


def check_correlator_products(P, orc, ctx, ham):
    """dmrgx_hshell_create_product (CalculateOperatorProducts + KronConstruct, include/DMRGBlockContainer.hpp:2262-2296,
    2340-2425): <psi| (O_1 O_2 ...)_sys ⊗ (O'_1 ...)_env |psi> against dense numpy products of the oracle's operators."""
    import ctypes as C
    SM = -1  # Op_t OpSm (include/DMRGBlock.hpp:21-27)

    d = orc.DMRG(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"])
    d.warmup(20)
    oL = orc.kron_eye(d.block(6), orc.Block.single_site(), lr_terms(orc, ham, 8))
    okb = orc.KronBlocks(oL, oL, [0.0])
    pL = upload_block(P, ctx, oL, orc)
    pkb = P.KronBlocks(pL, pL, [0.0])
    rng = np.random.default_rng(5)
    psi = rng.standard_normal(okb.num_states()); psi /= np.linalg.norm(psi)
    dpsi = ctx.vec(len(psi), psi)
    qn, il, ir, sz, off = okb.data()
    oqn, osz = oL.sectors()
    ooff = np.concatenate([[0], np.cumsum(osz)])
    dense = {(op, i): oL.get_op_dense(op, i) for op in (orc.OP_SZ, orc.OP_SP) for i in range(8)}
    for i in range(8):
        dense[(SM, i)] = dense[(orc.OP_SP, i)].T

    def expect_dense(lops, rops):
        n = oL.nstates
        A = np.eye(n); B = np.eye(n)
        for o in lops:
            A = A @ dense[o]
        for o in rops:
            B = B @ dense[o]
        # psi as a block matrix over sector pairs: Psi[l, r] with l, r global block indices
        Psi = np.zeros((n, n))
        for p in range(len(qn)):
            nl, nr = osz[il[p]], osz[ir[p]]
            Psi[ooff[il[p]]:ooff[il[p]] + nl, ooff[ir[p]]:ooff[ir[p]] + nr] = psi[off[p]:off[p + 1]].reshape(nl, nr)
        return float(np.sum(Psi * (A @ Psi @ B.T)))

    L = P.lib()
    cases = [([(orc.OP_SZ, 3)], []), ([(orc.OP_SZ, 2), (orc.OP_SZ, 5)], []), ([(orc.OP_SP, 1), (SM, 6)], []),
             ([(orc.OP_SZ, 7)], [(orc.OP_SZ, 7)]), ([(orc.OP_SP, 4)], [(SM, 0), (orc.OP_SZ, 3)]),
             ([(orc.OP_SZ, 0), (orc.OP_SZ, 1), (orc.OP_SZ, 2)], [(orc.OP_SZ, 6), (orc.OP_SZ, 5)]), ([], [(orc.OP_SZ, 2)])]
    for lops, rops in cases:
        h = C.c_void_p()
        la = np.array([o for o, _ in lops], np.int32); ls = np.array([s for _, s in lops], np.int64)
        ra = np.array([o for o, _ in rops], np.int32); rs = np.array([s for _, s in rops], np.int64)
        e = L.dmrgx_hshell_create_product(pkb.h, C.c_longlong(len(lops)), la.ctypes.data_as(C.c_void_p), ls.ctypes.data_as(C.c_void_p),
                                          C.c_longlong(len(rops)), ra.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p), C.byref(h))
        assert e == 0, L.dmrgx_last_error()
        v = C.c_double()
        assert L.dmrgx_expect(h, dpsi.ptr, C.byref(v)) == 0
        L.dmrgx_hshell_destroy(h)
        ref = expect_dense(lops, rops)
        assert abs(v.value - ref) < 1e-12, (lops, rops, v.value, ref)


def check_testkron01_kronblocks(P, ctx, golden_dir):
    """KronBlocks_t bookkeeping of the product against the reference's own golden case (tests/UnitTests_DMRGKron.cpp:39-252,
    TestKron01): left sectors {+.5:2, -.5:1}, right sectors {+1:1, 0:2, -1:1}, all sectors kept (IL-major, stable sort by
    descending total QN, include/DMRGKron.hpp:147-158)."""
    import json
    import os
    fx = json.load(open(os.path.join(golden_dir, "testkron01.json")))
    L = P.Block.Initialize(ctx, fx["blocks"]["Left"]["nsites"], fx["blocks"]["Left"]["qn"], fx["blocks"]["Left"]["sizes"])
    R = P.Block.Initialize(ctx, fx["blocks"]["Right"]["nsites"], fx["blocks"]["Right"]["qn"], fx["blocks"]["Right"]["sizes"])
    kb = P.KronBlocks(L, R, [])
    q, il, ir, size, off = kb.data()
    assert list(zip(q.tolist(), il.tolist(), ir.tolist(), size.tolist())) == [
        (1.5, 0, 0, 2), (0.5, 0, 1, 4), (0.5, 1, 0, 1), (-0.5, 0, 2, 2), (-0.5, 1, 1, 2), (-1.5, 1, 2, 1)]
    assert off.tolist() == [0, 2, 6, 7, 9, 11, 12] and kb.NumStates() == 12 and kb.size() == 6
    assert kb.Map(1, 1) == 4 and kb.Map(2, 0) == -1 and kb.Offsets(5, 5) == -1 and kb.Offsets(0, 1) == 2
    # one target sector: IL-major order, not sorted (include/DMRGKron.hpp:160-171)
    k0 = P.KronBlocks(L, R, [0.5])
    q, il, ir, size, off = k0.data()
    assert list(zip(il.tolist(), ir.tolist(), size.tolist())) == [(0, 1, 4), (1, 0, 1)] and off.tolist() == [0, 4, 5]


def check_step_fixture(P, ctx, golden_dir):
    """The product against the frozen step of tests/golden/step_j1j2_4x4.json (made by tests/golden/make_step_golden.py with the
    oracle): bookkeeping bit-exact, H·x to 1e-13, E0 and truncation errors to 1e-10, kept-state counts exact."""
    import json
    import os
    fx = json.load(open(os.path.join(golden_dir, "step_j1j2_4x4.json")))
    ham = fx["ham"]
    blk = P.Block.Initialize(ctx, fx["nsites"], fx["qn"], fx["sizes"])
    for i in range(fx["nsites"]):
        for name, code in (("Sz", P.OpSz), ("Sp", P.OpSp)):
            o = fx["ops"]["%s%d" % (name, i)]
            blk.set_operator(code, i, o["rowptr"], o["col"], o["val"])
    o = fx["ops"]["H"]
    blk.set_operator(P.OpH, 0, o["rowptr"], o["col"], o["val"])
    T = lambda n: P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, ham["bcx"], ham["bcy"])
    enl = P.KronEye_Explicit(blk, P.Block.SingleSite(ctx), T(8))
    q, s = enl.sectors()
    assert q.tolist() == fx["enlarged"]["qn"] and s.tolist() == fx["enlarged"]["sizes"]
    kb = P.KronBlocks(enl, enl, [0.0])
    kq, il, ir, size, off = kb.data()
    k = fx["kron"]
    assert (kq.tolist(), il.tolist(), ir.tolist(), size.tolist(), off.tolist()) == (k["qn"], k["il"], k["ir"], k["size"], k["off"])
    H = kb.KronSumConstruct(T(16))
    x = np.array(fx["x"]); y_ref = np.array(fx["y"])
    y = H.MatMult_host(x)
    assert np.abs(y - y_ref).max() <= 1e-13 * np.abs(y_ref).max() * H.stats()["nterms"]
    e0, psi, st = H.EPSSolve(tol=1e-12)
    assert st["converged"] and abs(e0 - fx["e0"]) <= ENERGY_RTOL * abs(fx["e0"])
    assert abs(abs(psi.get() @ np.array(fx["psi"])) - 1.0) < 1e-8
    btL, btR = P.GetTruncation(kb, ctx.vec(len(fx["psi"]), fx["psi"]), fx["mstates"])
    for bt, ref in ((btL, fx["trunc"]["L"]), (btR, fx["trunc"]["R"])):
        bq, bs = bt.sectors()
        assert bq.tolist() == ref["qn"] and bs.tolist() == ref["sizes"]
        assert abs(bt.TruncErr - ref["err"]) <= ENERGY_RTOL * max(abs(ref["err"]), 1e-4)
        ev = np.sort(bt.spectrum()[0])[::-1]
        assert np.abs(ev - np.array(ref["spectrum"])).max() <= 1e-12


def check_testkron01_operator_rows(P, ctx, golden_dir):
    """The reference's only operator-level golden (tests/UnitTests_DMRGKron.cpp:39-252, TestKron01: a 3-site left block with
    sectors {+.5: 2, -.5: 1} times a 2-site right block with sectors {+1: 1, 0: 2, -1: 1}, matrix values == column indices,
    tests/UnitTests_Misc.cpp:15-18) replayed on the PRODUCT through the C ABI: all 120 expected rows of the 10 enlarged
    operators.  A stored 0.0 (the fixture's column 0) is indistinguishable from an absent entry in a device tile, so rows are
    compared on their non-zero entries."""
    import json
    import os
    fx = json.load(open(os.path.join(golden_dir, "testkron01.json")))

    def block(spec):
        blk = P.Block.Initialize(ctx, spec["nsites"], spec["qn"], spec["sizes"])
        n = int(sum(spec["sizes"]))
        for op, code in (("Sz", P.OpSz), ("Sp", P.OpSp)):
            for site in range(spec["nsites"]):
                rows = [[] for _ in range(n)]
                for r in spec["rows"]:
                    if r["op"] == op and r["site"] == site:
                        rows[r["row"]] = list(r["cols"])
                rowptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
                col = np.array([c for r in rows for c in r], np.int64)
                blk.set_operator(code, site, rowptr, col, col.astype(float))
        blk.set_operator(P.OpH, 0, np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0))
        return blk
    L, R = block(fx["blocks"]["Left"]), block(fx["blocks"]["Right"])
    out = P.KronEye_Explicit(L, R, [])
    assert out.NumSites() == 5 and out.NumStates() == 12 and out.CheckOperatorBlocks() == 0
    q, s = out.sectors()
    assert q.tolist() == [1.5, 0.5, -0.5, -1.5] and s.tolist() == [2, 5, 4, 1]
    ops, seen = {}, 0
    for e in fx["expected"]:
        key = (e["op"], e["site"])
        if key not in ops:
            ops[key] = out.get_operator(P.OpSz if e["op"] == "Sz" else P.OpSp, e["site"])
        rowptr, col, val = ops[key]
        r = e["row"]
        got = [(int(c), float(v)) for c, v in zip(col[rowptr[r]:rowptr[r + 1]], val[rowptr[r]:rowptr[r + 1]]) if v != 0.0]
        exp = [(int(c), float(v)) for c, v in zip(e["cols"], e["vals"]) if v != 0.0]
        assert got == exp, (key, r, got, exp)
        seen += 1
    assert len(ops) == 10 and seen == len(fx["expected"]) == 120


def run_sweeps_with_prediction(P, ctx, ham, m_list, tol=1e-10, use_guess=True):
    """The reference's warm-up and sweep schedule (include/DMRGBlockContainer.hpp:809-840, 996-1088) on the product, with the
    wave-function transformation of csrc/predict.cpp evaluated at every step: returns one record per sweep step
    {guess: bool, overlap, nmatvec, energy}.  `use_guess=False` still computes the prediction (for the overlap) but starts the
    eigen-solve from the random vector, as the reference does."""
    N = ham["Lx"] * ham["Ly"]
    def terms(n):
        return P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, ham.get("bcx", 0), ham.get("bcy", 1))
    site = P.Block.SingleSite(ctx)
    serial = [0]
    blocks, bts = {}, {}          # index -> (Block, serial) ; index -> (BasisTransformation, serial of the block it enlarged)
    def put(idx, blk, bt, src_serial):
        serial[0] += 1
        blocks[idx] = (blk, serial[0])
        bts[idx] = (bt, src_serial)
        return serial[0]
    blocks[0] = (site, 0)
    pending = {}
    records = []

    def step(insys, inenv, outsys, outenv, m, in_sweep):
        (bs, ss), (be, se) = blocks[insys], blocks[inenv]
        L = P.KronEye_Explicit(bs, site, terms(insys + 2))
        R = P.KronEye_Explicit(be, site, terms(inenv + 2))
        kb = P.KronBlocks(L, R, [0.0])
        sh = kb.KronSumConstruct(terms(insys + inenv + 4))
        guess = None
        if in_sweep and pending:
            if pending["out_sys"] == ss and pending["xe_left"] is not None and pending["xe_left"][1] == se:
                guess = pending["wL"].apply(pending["xe_left"][0], site, kb)
            elif pending["out_env"] == se and pending["xe_right"] is not None and pending["xe_right"][1] == ss:
                guess = pending["wR"].apply(pending["xe_right"][0], site, kb)
        e, psi, st = sh.EPSSolve(tol=tol, initial=guess if use_guess else None)
        assert st["converged"]
        if in_sweep:
            rec = dict(guess=guess is not None, nmatvec=st["nmatvec"], energy=e, overlap=None)
            if guess is not None:
                g = guess.get(); x = psi.get()
                rec["overlap"] = abs(g @ x) / np.linalg.norm(g)
            records.append(rec)
        btL, btR = P.GetTruncation(kb, psi, m)
        xe_left, xe_right = bts.get(inenv), bts.get(insys)   # captured BEFORE this step's outputs replace anything
        wL, wR = P.Wave(kb, psi, btL, True), P.Wave(kb, psi, btR, False)
        new_sys = P.RotateOperators(L, btL)
        s_sys = put(outsys, new_sys, btL, ss)
        s_env = -1
        if outsys != outenv:
            new_env = P.RotateOperators(R, btR)
            s_env = put(outenv, new_env, btR, se)
        pending.clear()
        pending.update(wL=wL, wR=wR, out_sys=s_sys, out_env=s_env, xe_left=xe_left, xe_right=xe_right)

    for n in range(1, N // 2):                                  # warm-up: both blocks grow
        step(n - 1, n - 1, n, n, m_list[0], False)
    for m in m_list:
        pending.clear()
        for ib in range(N // 2, N - 3):
            step(ib - 1, N - ib - 3, ib, N - ib - 2, m, True)
        for ib in range(1, N // 2):
            step(N - ib - 3, ib - 1, N - ib - 2, ib, m, True)
    return records


def check_wavefunction_prediction(P, ctx):
    """(1) un-truncated blocks: the transformed vector IS the next ground state (overlap 1): pins every index map of
    csrc/predict.cpp, both growth directions; (2) truncating sweeps: same energies with and without the guess, fewer H*psi."""
    ham = dict(Lx=8, Ly=1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=0)
    recs = run_sweeps_with_prediction(P, ctx, ham, [64], tol=1e-12)
    with_guess = [r for r in recs if r["guess"]]
    assert len(with_guess) >= len(recs) - 3, [r["guess"] for r in recs]
    for r in with_guess:
        assert r["overlap"] > 1.0 - 1e-9, recs
        assert abs(r["energy"] - (-3.374932598688)) < 1e-9
    ham = dict(Lx=12, Ly=1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=0)
    a = run_sweeps_with_prediction(P, ctx, ham, [10, 10], tol=1e-10, use_guess=True)
    b = run_sweeps_with_prediction(P, ctx, ham, [10, 10], tol=1e-10, use_guess=False)
    assert [r["guess"] for r in a] == [r["guess"] for r in b]
    for ra, rb in zip(a, b):
        assert abs(ra["energy"] - rb["energy"]) < 1e-7, (ra, rb)
    na = sum(r["nmatvec"] for r in a if r["guess"]); nb = sum(r["nmatvec"] for r in b if r["guess"])
    assert na < 0.7 * nb, (na, nb)
    assert min(r["overlap"] for r in a if r["guess"]) > 0.9
    return recs, a, b


def check_wavefunction_prediction_rejects_mismatch(P, ctx):
    """dmrgx_wave_apply must answer ok = 0 (no exception, nothing written that a caller would use) when the transformation
    handed in does not chain the previous superblock to the new one, and dmrgx_wave_create must refuse a transformation of
    the wrong side."""
    ham = dict(Lx=8, Ly=1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=0)
    def terms(n):
        return P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, 0, 0)
    site = P.Block.SingleSite(ctx)
    b1 = P.KronEye_Explicit(site, site, terms(2))            # two sites, exact
    kb22 = P.KronBlocks(P.KronEye_Explicit(b1, site, terms(3)), P.KronEye_Explicit(b1, site, terms(3)), [0.0])   # 3 + 3 sites
    e, psi, st = kb22.KronSumConstruct(terms(6)).EPSSolve(tol=1e-12)
    btL, btR = P.GetTruncation(kb22, psi, 6)                  # truncating: 8 -> 6 states
    w = P.Wave(kb22, psi, btL, True)
    # a superblock that does not follow from this one: (2 + 1) + (1 + 1) sites
    kb_other = P.KronBlocks(P.KronEye_Explicit(b1, site, terms(3)), P.KronEye_Explicit(site, site, terms(2)), [0.5])
    assert w.apply(btR, site, kb_other) is None
    assert w.apply(btL, site, kb22) is None
    # the left transformation of an asymmetric superblock is not a transformation of its right block
    kb32 = P.KronBlocks(P.KronEye_Explicit(b1, site, terms(3)), P.KronEye_Explicit(site, site, terms(2)), [0.5])
    e2, psi2, _ = kb32.KronSumConstruct(terms(5)).EPSSolve(tol=1e-12)
    bl, br = P.GetTruncation(kb32, psi2, 4)
    try:
        P.Wave(kb32, psi2, bl, False)
        raise AssertionError("wave_create accepted the transformation of the other side")
    except P.DmrgxError as exc:
        assert exc.code == 62
