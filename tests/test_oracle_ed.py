"""Pins the oracle's arithmetic (matvec, Lanczos, truncation, rotation, enlargement, bond rules) against
exact diagonalisation: the reference holds no test for these, so the known answers are
regenerated here with scipy on the full 2^N Hilbert space restricted to Sz=0, using an independent
statement of the bond rules of src/Hamiltonians.cpp:26-122.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def snake(ix, jy, Ly):
    return (ix * Ly + jy) if ix % 2 == 0 else ((ix + 1) * Ly - (jy + 1))


def bonds(Lx, Ly, bcx, bcy, nnn):
    """independent enumeration: every site looks 'up' and 'right' (and diagonally up-left/up-right)"""
    out = []
    for ix in range(Lx):
        for jy in range(Ly):
            s = snake(ix, jy, Ly)
            if jy < Ly - 1 or bcy:
                j2 = (jy + 1) % Ly
                if j2 != jy:
                    out.append((1, s, snake(ix, j2, Ly)))
            if ix < Lx - 1 or bcx:
                i2 = (ix + 1) % Lx
                if i2 != ix:
                    out.append((1, s, snake(i2, jy, Ly)))
            if nnn and Lx > 1 and Ly > 1:
                if (ix >= 1 or bcx) and (jy < Ly - 1 or bcy):
                    out.append((2, s, snake((ix + Lx - 1) % Lx, (jy + 1) % Ly, Ly)))
                if (ix < Lx - 1 or bcx) and (jy < Ly - 1 or bcy):
                    out.append((2, s, snake((ix + 1) % Lx, (jy + 1) % Ly, Ly)))
    return out


def ed_energy(Lx, Ly, J1, Jz1, J2, Jz2, bcx=0, bcy=1):
    N = Lx * Ly
    states = np.array([s for s in range(1 << N) if bin(s).count("1") == N // 2], dtype=np.int64)
    index = {int(s): i for i, s in enumerate(states)}
    rows, cols, vals = [], [], []
    nnn = (J2 != 0.0 and Jz2 != 0.0)  # reference quirk, src/Hamiltonians.cpp:101
    for kind, a, b in bonds(Lx, Ly, bcx, bcy, nnn):
        J, Jz = (J1, Jz1) if kind == 1 else (J2, Jz2)
        ba = (states >> a) & 1
        bb = (states >> b) & 1
        diag = Jz * (ba - 0.5) * (bb - 0.5)
        rows.extend(range(len(states))); cols.extend(range(len(states))); vals.extend(diag.tolist())
        if J != 0.0:
            flip = np.nonzero(ba != bb)[0]
            tgt = states[flip] ^ ((1 << a) | (1 << b))
            rows.extend(flip.tolist()); cols.extend(index[int(t)] for t in tgt); vals.extend([J] * len(flip))
    H = sp.csr_matrix((vals, (rows, cols)), shape=(len(states), len(states)))
    if H.shape[0] < 200:
        return float(np.linalg.eigvalsh(H.toarray())[0])
    return float(spla.eigsh(H, k=1, which="SA", tol=1e-13)[0][0])


def dmrg_mid_energy(O, m, sweeps, **kw):
    d = O.DMRG(**kw)
    d.warmup(m)
    for _ in range(sweeps):
        d.sweep(m)
    N = kw["Lx"] * kw["Ly"]
    mid = [s for s in d.steps() if s["NSites_Sys"] == s["NSites_Env"] and s["NSites_SysEnl"] + s["NSites_EnvEnl"] == N]
    return mid[-1]["GSEnergy"], d.steps()


CASES = [
    # name, kwargs, m, sweeps, BASELINE.md §3 known answer
    ("chain8", dict(Lx=8, Ly=1, heisenberg=1.0, bcx=0, bcy=0), 16, 1, -3.374932598688),
    ("chain12", dict(Lx=12, Ly=1, heisenberg=1.0, bcx=0, bcy=0), 64, 1, -5.142090632841),
    ("4x2heis", dict(Lx=4, Ly=2, heisenberg=1.0), 16, 1, -6.668276634635),
]


@pytest.mark.parametrize("name,kw,m,sweeps,known", CASES, ids=[c[0] for c in CASES])
def test_oracle_energy_vs_ed(orc, name, kw, m, sweeps, known):
    e, steps = dmrg_mid_energy(orc, m, sweeps, **kw)
    p = dict(J1=1.0, Jz1=0.0, J2=1.0, Jz2=0.0, bcx=0, bcy=1)
    if "heisenberg" in kw:
        p.update(J1=0.5, Jz1=kw["heisenberg"], J2=0.0, Jz2=0.0)
    p.update({k: v for k, v in kw.items() if k in p})
    ed = ed_energy(kw["Lx"], kw["Ly"], p["J1"], p["Jz1"], p["J2"], p["Jz2"], p["bcx"], p["bcy"])
    assert abs(ed - known) < 1e-10 * abs(known)
    assert abs(e - ed) < 1e-10 * abs(ed), (e, ed)
    # exact runs: truncation errors are round-off only
    assert all(abs(s["TruncErr_Sys"]) < 1e-12 for s in steps)


def test_oracle_energy_vs_ed_nnn(orc):
    """Cylinders with NNN terms active (Jz2 != 0): next-nearest-neighbour bond rules incl. the Ly=2 wrap."""
    for Lx, Ly, m in ((4, 2, 16), (6, 2, 64), (3, 4, 64)):
        kw = dict(Lx=Lx, Ly=Ly, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5)
        e, _ = dmrg_mid_energy(orc, m, 1, **kw)
        ed = ed_energy(Lx, Ly, 0.5, 1.0, 0.25, 0.5)
        assert abs(e - ed) < 1e-10 * abs(ed), (Lx, Ly, e, ed)


def test_ed_reproduces_baseline_4x4_values():
    """The independent ED statement reproduces BASELINE.md §3 (4x4 cylinder rows), so the bond rules agree."""
    assert abs(ed_energy(4, 4, 1.0, 0.0, 1.0, 0.0) - (-16.033548229533)) < 1e-9
    assert abs(ed_energy(4, 4, 0.5, 1.0, 0.0, 0.0) - (-10.264289620979)) < 1e-9
    assert abs(ed_energy(4, 4, 0.5, 1.0, 0.25, 0.5) - (-8.261232563030)) < 1e-9


def test_term_counts_match_survey_table(orc):
    """L-R term counts at the midpoint cut."""
    O = orc

    def lr(Lx, Ly, **kw):
        p = dict(J1=1.0, Jz1=0.0, J2=1.0, Jz2=0.0); p.update(kw)
        N = Lx * Ly
        t = O.ham_terms(Lx, Ly, p["J1"], p["Jz1"], p["J2"], p["Jz2"], N, kw.get("bcx", 0), kw.get("bcy", 1))
        return sum(1 for x in t if x[2] < N // 2 <= x[4])

    assert lr(24, 1, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0, bcy=0) == 3
    assert lr(8, 4, J1=0.5, Jz1=1.0, J2=0.0, Jz2=0.0) == 12
    assert lr(12, 6, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5) == 54
    assert lr(12, 6, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.0) == 18  # NNN dropped by the quirk at src/Hamiltonians.cpp:101
    assert lr(16, 8, J1=1.0, Jz1=0.0, J2=1.0, Jz2=0.0) == 16


def test_truncated_run_is_variational_and_converges(orc):
    """m below the exact dimension: energies stay above ED and approach it with sweeps."""
    kw = dict(Lx=12, Ly=1, heisenberg=1.0, bcx=0, bcy=0)
    e8, steps = dmrg_mid_energy(orc, 8, 2, **kw)
    e16, _ = dmrg_mid_energy(orc, 16, 2, **kw)
    ed = -5.142090632841
    assert e8 > ed - 1e-12 and e16 > ed - 1e-12
    assert abs(e16 - ed) < abs(e8 - ed) < 1e-3
    assert max(s["NStates_SysRot"] for s in steps) <= 8
