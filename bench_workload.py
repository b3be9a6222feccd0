"""Synthetic superblock workloads in the reference's block layout (BASELINE.json configs[4], SURVEY.md §8d "C5").

A workload is the pair of (un-enlarged) system / environment blocks a sweep-midpoint SingleDMRGStep of the named
lattice would see after truncation to m states: sector sizes n(k) = round(m·G(k; sigma=2.5)), operators dense
N(0,1)/sqrt(n) inside their allowed sector blocks (what rotation by a dense U produces), H symmetric; the added
site carries the exact single-site operators.  The data is produced on the HOST as CSR with global column indices —
exactly what a maintainer of the reference would hand over from MatGetRow — and is given unchanged to the product
(through dmrgx_block_set_operator) and, for parity tests and the CPU baseline only, to the oracle.
Only the operators the Hamiltonian terms of the step touch are filled; the other sites get empty operators.
"""
import os

import numpy as np

SEED = 20261018

CONFIGS = {
    # BASELINE.json configs[2] / SURVEY §8d C3: -Lx 12 -Ly 6 -J1 0.5 -Jz1 1 -J2 0.25 -Jz2 0.5 (J2/J1 = 0.5, NNN terms active)
    "j1j2_12x6": dict(Lx=12, Ly=6, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5, bcx=0, bcy=1),
    # configs[1] / C2: -Lx 8 -Ly 4 -heisenberg 1
    "heis_8x4": dict(Lx=8, Ly=4, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=1),
    # configs[0] / C1: -Lx 24 -Ly 1 -heisenberg 1 -BCopen
    "heis_chain24": dict(Lx=24, Ly=1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcx=0, bcy=0),
    # configs[3] / C4: XY model on 16x8 (reference semantics: NNN dropped because Jz2 == 0)
    "xy_16x8": dict(Lx=16, Ly=8, J1=1.0, Jz1=0.0, J2=1.0, Jz2=0.0, bcx=0, bcy=1),
}


def sector_sizes(m, nsites, sigma=2.5):
    """Gaussian sector model: descending QN list with unit steps, sizes summing to m."""
    half = (nsites % 2 == 1)
    ks = np.arange(-12, 13)
    qn = ks + (0.5 if half else 0.0)
    w = np.exp(-0.5 * (qn / sigma) ** 2)
    sz = np.floor(m * w / w.sum() + 0.5).astype(np.int64)
    # keep a contiguous run of non-empty sectors, fix the total on the central sectors
    keep = np.nonzero(sz > 0)[0]
    qn, sz = qn[keep[0]:keep[-1] + 1], sz[keep[0]:keep[-1] + 1]
    c = len(sz) // 2
    sz[c] += m - sz.sum()
    rnd = int(os.environ.get("DMRGX_SECTOR_ROUND", "0"))   # experiment hook: sector sizes in multiples of `rnd` (no ragged tiles)
    if rnd > 1:
        sz = np.maximum(rnd, (sz + rnd // 2) // rnd * rnd)
    order = np.argsort(-qn)
    return qn[order].tolist(), sz[order].tolist()


def _dense_sector_csr(rng, sizes, shift, symmetric=False):
    """CSR (global indices) of an operator that is dense inside the sector blocks (I, I+shift)."""
    off = np.concatenate([[0], np.cumsum(sizes)])
    n = int(off[-1])
    rowlen = np.zeros(n, np.int64)
    cols, vals = [], []
    for I in range(len(sizes)):
        J = I + shift
        if J < 0 or J >= len(sizes) or sizes[I] == 0 or sizes[J] == 0:
            continue
        blk = rng.standard_normal((sizes[I], sizes[J])) / np.sqrt(max(sizes[I], sizes[J]))
        if symmetric:
            blk = 0.5 * (blk + blk.T)
        rowlen[off[I]:off[I + 1]] = sizes[J]
        cols.append(np.tile(np.arange(off[J], off[J + 1]), sizes[I]))
        vals.append(blk.ravel())
    rowptr = np.concatenate([[0], np.cumsum(rowlen)])
    if cols:
        return rowptr, np.concatenate(cols), np.concatenate(vals)
    return rowptr, np.zeros(0, np.int64), np.zeros(0)


def used_sites(terms_enlarge, terms_super, nsites_blk):
    """Sites of the un-enlarged block whose Sz/Sp enter (a) its enlargement and (b) the L-R shell terms."""
    used = set()
    nenl = nsites_blk + 1
    for (_, _, i, _, j) in terms_enlarge:
        if i < nsites_blk <= j < nenl:
            used.add(i)
    ntot = 2 * nenl
    for (_, _, i, _, j) in terms_super:
        if i < nenl <= j < ntot:
            if i < nsites_blk:
                used.add(i)
            jr = ntot - 1 - j  # reflection, src/DMRGKron.cpp:803-807 (the environment is the mirrored system block)
            if jr < nsites_blk:
                used.add(jr)
    return sorted(used)


def synth_block_host(m, nsites, used, seed=SEED, sigma=2.5):
    """Host-side description of one truncated block: sector lists + CSR of Sz_i / Sp_i (i in `used`) and H."""
    rng = np.random.default_rng(seed)
    qn, sz = sector_sizes(m, nsites, sigma)
    ops = {}
    n = int(np.sum(sz))
    empty = (np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0))
    for i in range(nsites):
        if i in used:
            ops[("Sz", i)] = _dense_sector_csr(rng, sz, 0, symmetric=True)
            ops[("Sp", i)] = _dense_sector_csr(rng, sz, +1)
        else:
            ops[("Sz", i)] = empty
            ops[("Sp", i)] = empty
    ops["H"] = _dense_sector_csr(rng, sz, 0, symmetric=True)
    return dict(nsites=nsites, qn=qn, sizes=sz, ops=ops)


def algorithmic_bytes(host, terms_super, nsites_blk, D):
    """SURVEY.md §8d, without touching the library: 16*D (psi in, y out) + every distinct ORIGINAL operator panel once —
    Sz_i / Sp_i of the block sites that appear in L-R terms (the added site's operators are scaled identities, 0 bytes; the
    environment is the mirrored system block, so both sides share the panels) + the enlarged block's H, dense per sector."""
    nenl = nsites_blk + 1
    sites = set()
    for (a, _, i, _, j) in terms_super:
        if a != 0 and i < nenl <= j < 2 * nenl:
            sites.add(i); sites.add(2 * nenl - 1 - j)
    tb = 0
    for i in sites:
        if i < nsites_blk:
            tb += 8 * len(host["ops"][("Sz", i)][2]) + 8 * len(host["ops"][("Sp", i)][2])
    enl = {}
    for q, n in zip(host["qn"], host["sizes"]):
        for dq in (+0.5, -0.5):
            enl[q + dq] = enl.get(q + dq, 0) + int(n)
    tb += 8 * sum(v * v for v in enl.values())
    return 16 * D + tb


class Workload:
    """Everything one H·psi benchmark / parity case needs, on the product side."""

    def __init__(self, P, ctx, config="j1j2_12x6", m=2048, seed=SEED):
        self.P, self.ctx, self.config, self.m = P, ctx, config, m
        ham = CONFIGS[config]
        self.ham = ham
        N = ham["Lx"] * ham["Ly"]
        self.nsites_blk = N // 2 - 1
        self.terms_enl = P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], self.nsites_blk + 1,
                                            ham["bcx"], ham["bcy"])
        self.terms = P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], N, ham["bcx"], ham["bcy"])
        self.used = used_sites(self.terms_enl, self.terms, self.nsites_blk)
        self.host = synth_block_host(m, self.nsites_blk, self.used, seed)
        blk = P.Block.Initialize(ctx, self.nsites_blk, self.host["qn"], self.host["sizes"])
        for i in range(self.nsites_blk):
            blk.set_operator(P.OpSz, i, *self.host["ops"][("Sz", i)])
            blk.set_operator(P.OpSp, i, *self.host["ops"][("Sp", i)])
        blk.set_operator(P.OpH, 0, *self.host["ops"]["H"])
        self.blk = blk
        self.site = P.Block.SingleSite(ctx)
        # SingleDMRGStep with &SysBlock == &EnvBlock (include/DMRGBlockContainer.hpp:1327-1343): one enlargement
        self.enl = P.KronEye_Explicit(blk, self.site, self.terms_enl)
        self.kron = P.KronBlocks(self.enl, self.enl, [0.0])
        self.shell = self.kron.KronSumConstruct(self.terms)
        self.n = self.kron.NumStates()

    def random_state(self, seed=1):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(self.n)
        return x / np.linalg.norm(x)


def oracle_side(O, wl):
    """The same inputs on the oracle (parity tests / CPU baseline only): returns (enlarged block, KronBlocks)."""
    h = wl.host
    b = O.Block.create(h["nsites"], h["qn"], h["sizes"])
    for i in range(h["nsites"]):
        b.set_op(O.OP_SZ, i, *h["ops"][("Sz", i)])
        b.set_op(O.OP_SP, i, *h["ops"][("Sp", i)])
    b.set_op(O.OP_H, 0, *h["ops"]["H"])
    enl = O.kron_eye(b, O.Block.single_site(), wl.terms_enl)
    kb = O.KronBlocks(enl, enl, [0.0])
    return enl, kb


class ExactChainWorkload:
    """The sparse-sector case of the north star: UN-truncated blocks.  An open Heisenberg chain of 2*nhalf sites cut in the
    middle, both halves represented exactly (2^nhalf states, operators with O(1) non-zeros per row).  The operators are
    handed over as CSR with global column indices — what a maintainer of the reference passes from MatGetRow (route B of
    INTEGRATION.md) — so the library stores them as CSR / scaled-identity tiles and H*psi runs on the sparse segments of
    the chain kernel instead of the FP64 tensor path.  Bound: HBM."""

    def __init__(self, P, ctx, nhalf=12):
        ham = CONFIGS["heis_chain24"]
        self.P, self.ctx, self.nhalf = P, ctx, nhalf
        Lx = 2 * nhalf
        T = lambda n: P.HamiltonianTerms(Lx, 1, ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, ham["bcx"], ham["bcy"])
        site = P.Block.SingleSite(ctx)
        blk = P.Block.SingleSite(ctx)
        for n in range(2, nhalf + 1):
            blk = P.KronEye_Explicit(blk, site, T(n))
        # round trip through host CSR: the upload path classifies every sector block by fill (dense / CSR / identity runs)
        qn, sz = blk.sectors()
        ctx.set_dense_threshold(0.125)
        up = P.Block.Initialize(ctx, nhalf, qn, sz)
        nnz = 0
        for i in range(nhalf):
            for op in (P.OpSz, P.OpSp):
                rp, ci, vv = blk.get_operator(op, i)
                up.set_operator(op, i, rp, ci, vv)
        rp, ci, vv = blk.get_operator(P.OpH, 0)
        nnz = len(vv)
        up.set_operator(P.OpH, 0, rp, ci, vv)
        self.h_nnz_per_row = nnz / float(up.NumStates())
        self.blk = up
        self.terms = T(Lx)
        self.kron = P.KronBlocks(up, up, [0.0])
        self.shell = self.kron.KronSumConstruct(self.terms)
        self.n = self.kron.NumStates()

    def random_state(self, seed=1):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(self.n)
        return x / np.linalg.norm(x)


class DiskWorkload(Workload):
    """BASELINE.json configs[4]: the isolated H*psi on blocks read from disk (InitializeFromDisk layout) — a Sweep_*/ directory
    written by the reference or by `DMRG-SquareLattice.x -scratch_dir`: the sweep-midpoint superblock of its Sys_{N/2-2} block."""

    def __init__(self, P, ctx, config, sweep_dir):
        self.P, self.ctx, self.config = P, ctx, config
        ham = CONFIGS[config]
        self.ham = ham
        N = ham["Lx"] * ham["Ly"]
        self.nsites_blk = N // 2 - 1
        T = lambda n: P.HamiltonianTerms(ham["Lx"], ham["Ly"], ham["J1"], ham["Jz1"], ham["J2"], ham["Jz2"], n, ham["bcx"], ham["bcy"])
        self.terms_enl, self.terms = T(self.nsites_blk + 1), T(N)
        self.blk = P.Block.InitializeFromDisk(ctx, os.path.join(sweep_dir, "Sys_%09d" % (self.nsites_blk - 1)))
        assert self.blk.NumSites() == self.nsites_blk
        self.m = self.blk.NumStates()
        self.site = P.Block.SingleSite(ctx)
        self.enl = P.KronEye_Explicit(self.blk, self.site, self.terms_enl)
        self.kron = P.KronBlocks(self.enl, self.enl, [0.0])
        self.shell = self.kron.KronSumConstruct(self.terms)
        self.n = self.kron.NumStates()
