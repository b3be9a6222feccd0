"""-m gpu: the real DMRG-SquareLattice.x (libdmrgx_b200.so, sm_100a kernels) against the oracle's DMRG loop and the
exact-diagonalisation energies of SURVEY.md §8c.  BASELINE.json configs[0]: Heisenberg chain Lx=24, mwarmup 32,
msweeps 64,128, E0 = -10.453785760410."""
import os

import pytest

import driver_common as dc

pytestmark = pytest.mark.gpu
EXE = os.path.join(dc.ROOT, "dmrg.x_b200", "DMRG-SquareLattice.x")


def test_config0_heisenberg_chain24(orc, tmp_path):
    assert os.path.exists(EXE), "build it with __graft_entry__.build()"
    docs, out = dc.run_driver(EXE, tmp_path, ["-Lx", 24, "-Ly", 1, "-heisenberg", 1, "-BCopen"], 32, [64, 128])
    ref, tie = dc.compare_with_oracle(orc, docs, dict(Lx=24, Ly=1, heisenberg=1.0, bcx=0, bcy=0), 32, [64, 128])
    hdr = docs["DMRGSteps"]["headers"]
    mid = [dict(zip(hdr, r)) for r in docs["DMRGSteps"]["table"] if r[hdr.index("NSites_Sys")] == r[hdr.index("NSites_Env")] == 11]
    e = mid[-1]["GSEnergy"]
    assert abs(e - (-10.453785760410)) < 1e-8, e                      # exact diagonalisation, Sz = 0 sector
    assert abs(e - ref[-1]["GSEnergy"]) <= 1e-10 * abs(e)             # and the oracle's DMRG at the same step
    assert "kernel launches" in out and int(out.rsplit("kernel launches:", 1)[1].split()[0]) > 0
    dc.check_correlations(docs, 24)


def test_j1j2_cylinder_6x4(orc, tmp_path):
    args = ["-Lx", 6, "-Ly", 4, "-J1", 0.5, "-Jz1", 1, "-J2", 0.25, "-Jz2", 0.5]
    docs, out = dc.run_driver(EXE, tmp_path, args, 32, [64])
    dc.compare_with_oracle(orc, docs, dict(Lx=6, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5), 32, [64])
