"""CPU-only checks of the product's HOST logic (tile classification, sector/offset maps, work-item and segment
lists, Lanczos restart logic, truncation selection): the real host code of dmrg.x_b200/csrc is linked against the
test-only emulation of the device layer (tests/plancheck/dev_host.cpp) and compared with the oracle.  The CUDA
kernels themselves are checked by tests/test_gpu_parity.py (-m gpu) through the same shared checks."""
import os
import subprocess

import numpy as np

import libswitch
import pytest

import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANCHECK = os.path.join(ROOT, "tests", "plancheck", "libdmrgx_plancheck.so")


@pytest.fixture(scope="module")
def P():
    import dmrgx_loader
    if not os.path.exists(PLANCHECK):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "dmrg.x_b200", "csrc"), "plancheck"])
    P = dmrgx_loader.load_package()
    libswitch.use_library(P, PLANCHECK)
    yield P
    libswitch.use_library(P, None)


@pytest.fixture()
def ctx(P):
    c = P.Context(0)
    yield c
    c.close()


HEIS_CHAIN = dict(Lx=12, Ly=1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=0)
J1J2_CYL = dict(Lx=4, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5)
HEIS_CYL = dict(Lx=6, Ly=2, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0)


def test_single_site_block(P, ctx, orc):
    for spin in (1, 2):
        b = P.Block.SingleSite(ctx, spin)
        o = orc.Block.single_site(spin)
        pc.assert_blocks_equal(P, orc, b, o, tol=0.0)


def test_upload_roundtrip_and_sector_violation(P, ctx, orc, golden_dir):
    import json
    fx = json.load(open(os.path.join(golden_dir, "opblocks.json")))
    b = P.Block.Initialize(ctx, 2, fx["sectors"]["qn"], fx["sectors"]["sizes"])
    for name, rowsspec in fx["check_cases"].items():
        op, site = name[:2], int(name[3])
        rowptr = [0]; col = []; val = []
        rows = {r["row"]: r["cols"] for r in rowsspec}
        for r in range(8):
            for c in rows.get(r, []):
                col.append(c); val.append(float(c) + 0.5)
            rowptr.append(len(col))
        code = P.OpSz if op == "Sz" else P.OpSp
        if fx["check_expect"][name] == 0:
            b.set_operator(code, site, rowptr, col, val)
            rp, ci, vv = b.get_operator(code, site)
            assert rp.tolist() == rowptr and ci.tolist() == col and vv.tolist() == val
        else:  # tests/UnitTests_DMRGBlock.cpp:112-128: PETSC_ERR_ARG_OUTOFRANGE
            with pytest.raises(P.DmrgxError) as e:
                b.set_operator(code, site, rowptr, col, val)
            assert e.value.code == fx["check_expect"][name] == 63


def test_kron_bookkeeping_all_sectors_testkron01(P, ctx, orc, golden_dir):
    """KronBlocks_t over the TestKron01 blocks, every sector kept: bit-exact order, offsets and map."""
    import json
    from test_oracle_golden import _block_from_fixture
    fx = json.load(open(os.path.join(golden_dir, "testkron01.json")))
    oL = _block_from_fixture(orc, fx["blocks"]["Left"]); oR = _block_from_fixture(orc, fx["blocks"]["Right"])
    pL = P.Block.Initialize(ctx, 3, fx["blocks"]["Left"]["qn"], fx["blocks"]["Left"]["sizes"])
    pR = P.Block.Initialize(ctx, 2, fx["blocks"]["Right"]["qn"], fx["blocks"]["Right"]["sizes"])
    for qn in ([], [0.5], [0.5, -0.5], [2.5], [7.5]):   # [7.5]: no such sector — an empty KronBlocks_t is legal in the reference (include/DMRGKron.hpp:160-171)
        pc.check_kron_bookkeeping(P, orc, P.KronBlocks(pL, pR, qn), orc.KronBlocks(oL, oR, qn))


@pytest.mark.parametrize("ham,nsys,nenv,mprep,mkeep", [
    (HEIS_CHAIN, 3, 3, 8, 6),      # truncating step on the chain, sys == env sizes
    (HEIS_CHAIN, 4, 2, 16, 6),     # asymmetric blocks (rank of rho_L limited by the small env)
    (HEIS_CYL, 4, 4, 12, 10),      # Ly=2 cylinder (doubled rung bonds), truncating
    (J1J2_CYL, 5, 5, 20, 16),      # NNN terms, rotated (dense) operators on both sides
], ids=["chain-3+3", "chain-4+2", "cyl6x2-4+4", "j1j2-4x4-5+5"])
def test_single_dmrg_step_matches_oracle(P, ctx, orc, ham, nsys, nenv, mprep, mkeep):
    rng = np.random.default_rng(7)
    pc.run_step_parity(P, orc, ctx, ham, nsys, nenv, mprep, mkeep, rng)


def test_sparse_and_dense_tile_paths_agree(P, ctx, orc):
    """Force every panel sparse (CSR) and every panel dense: H·psi must not depend on the storage choice."""
    rng = np.random.default_rng(11)
    ham = J1J2_CYL
    d = orc.DMRG(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"])
    d.warmup(12)
    osys = d.block(5); osite = orc.Block.single_site()
    oL = orc.kron_eye(osys, osite, pc.lr_terms(orc, ham, 7))
    okb = orc.KronBlocks(oL, oL, [0.0])
    terms = pc.lr_terms(orc, ham, 14)
    osh = orc.Shell(okb, terms)
    x = rng.standard_normal(osh.n)
    y_ref = osh.apply(x)
    for thr in (0.0, 2.0):  # 0: everything dense, 2: nothing dense
        ctx.set_dense_threshold(thr)
        pL = pc.upload_block(P, ctx, oL, orc)
        pkb = P.KronBlocks(pL, pL, [0.0])
        psh = pkb.KronSumConstruct(terms)
        y = psh.MatMult_host(x)
        assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    ctx.set_dense_threshold(0.125)


def test_correlator_single_term_shell(P, ctx, orc):
    """KronConstruct + MatMult + VecDot (include/DMRGBlockContainer.hpp:2287-2296)"""
    rng = np.random.default_rng(3)
    ham = HEIS_CHAIN
    d = orc.DMRG(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], 0, 0)
    d.warmup(8)
    oL = orc.kron_eye(d.block(2), orc.Block.single_site(), pc.lr_terms(orc, ham, 4))
    okb = orc.KronBlocks(oL, oL, [0.0])
    pL = pc.upload_block(P, ctx, oL, orc)
    pkb = P.KronBlocks(pL, pL, [0.0])
    x = rng.standard_normal(okb.num_states()); x /= np.linalg.norm(x)
    dx = ctx.vec(len(x), x)
    for (opl, il, opr, ir) in ((orc.OpSz, 1, orc.OpSz, 2), (orc.OpSp, 3, orc.OpSm, 0), (orc.OpSm, 0, orc.OpSp, 3)):
        osh = orc.Shell(okb, single=(opl, il, opr, ir))
        psh = pkb.KronConstruct(opl, il, opr, ir)
        ref = float(x @ osh.apply(x))
        assert abs(psh.expect(dx) - ref) < 1e-13


def test_no_device_no_fallback():
    """The product library refuses to run without a GPU (no CPU path)."""
    import ctypes as C
    import dmrgx_loader
    P = dmrgx_loader.load_package()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = C.CDLL(os.path.join(ROOT, "dmrg.x_b200", "libdmrgx_b200.so"))   # the PRODUCT library, not the emulation build
    h = C.c_void_p()
    assert lib.dmrgx_ctx_create(0, None, C.byref(h)) == 100


def test_exact_chain_sparse_workload_ground_state(P, ctx):
    """Un-truncated (sparse / identity-tile) blocks uploaded as CSR: the shell of the 12-site open Heisenberg chain cut in
    the middle must have the exact ground-state energy of SURVEY.md §8c."""
    import bench_workload as W
    sw = W.ExactChainWorkload(P, ctx, 6)
    assert sw.n == 924
    e, psi, st = sw.shell.EPSSolve(tol=1e-12)
    assert st["converged"] and abs(e - (-5.142090632841)) < 1e-9


def test_correlator_operator_products_match_dense(P, ctx, orc):
    pc.check_correlator_products(P, orc, ctx, J1J2_CYL)


def test_in_cycle_stopping_test_of_the_eigensolver(P, ctx, orc, monkeypatch):
    """The stopping test made inside a restart cycle (on by default only for large superblocks) gives the same eigenpair with
    fewer matvecs than testing at the restart boundaries only."""
    import bench_workload as W
    wl = W.Workload(P, ctx, "heis_8x4", m=48)
    monkeypatch.setenv("DMRGX_EARLY_TEST_MIN", "1000000000")
    e0, psi0, st0 = wl.shell.EPSSolve(tol=1e-10)
    monkeypatch.setenv("DMRGX_EARLY_TEST_MIN", "1")
    e1, psi1, st1 = wl.shell.EPSSolve(tol=1e-10)
    assert st0["converged"] and st1["converged"]
    assert abs(e0 - e1) <= 1e-9 * abs(e0) and abs(abs(psi0.get() @ psi1.get()) - 1.0) < 1e-6
    assert st1["nmatvec"] <= st0["nmatvec"] and st1["resid"] <= 1e-10 * abs(e1)
    x = psi1.get()
    assert abs(x @ wl.shell.MatMult_host(x) - e1) < 1e-9


def test_testkron01_kronblocks_golden(P, ctx, golden_dir):
    pc.check_testkron01_kronblocks(P, ctx, golden_dir)


def test_testkron01_operator_rows_on_the_product(P, ctx, golden_dir):
    pc.check_testkron01_operator_rows(P, ctx, golden_dir)


def test_frozen_step_fixture(P, ctx, golden_dir):
    pc.check_step_fixture(P, ctx, golden_dir)


def test_argument_validation_at_the_abi(P, ctx):
    """-H_eps_ncv beyond what the fused Lanczos kernels hold is refused with PETSC_ERR_ARG_OUTOFRANGE (63) and a message,
    instead of failing mid-solve; CSR row pointers that decrease are refused with PETSC_ERR_ARG_CORRUPT (64)."""
    import bench_workload as W
    sw = W.ExactChainWorkload(P, ctx, 4)
    with pytest.raises(P.DmrgxError) as e:
        sw.shell.EPSSolve(ncv=40)
    assert e.value.code == 63 and "ncv" in str(e.value)
    e0, _, st = sw.shell.EPSSolve(ncv=39, tol=1e-10)
    assert st["converged"]
    b = P.Block.Initialize(ctx, 1, [0.5, -0.5], [2, 2])
    with pytest.raises(P.DmrgxError) as e:
        b.set_operator(P.OpSz, 0, [0, 3, 1, 2, 2], [0, 1, 0], [1.0, 1.0, 1.0])
    assert e.value.code == 64
    with pytest.raises(P.DmrgxError) as e:
        b.set_operator(P.OpSz, 0, [1, 1, 1, 1, 1], [0], [1.0])
    assert e.value.code == 64


def test_wavefunction_prediction_extension(P, ctx):
    pc.check_wavefunction_prediction(P, ctx)
    pc.check_wavefunction_prediction_rejects_mismatch(P, ctx)
