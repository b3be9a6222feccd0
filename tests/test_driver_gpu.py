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


def test_restart_and_initialize_from_disk_on_the_device(tmp_path):
    """f-2 on the real executable: -scratch_dir checkpoints in the reference's on-disk format, -restart_dir continues from them
    and reproduces the uninterrupted run; the saved Sys block read back through Block.InitializeFromDisk (BASELINE configs[4])
    gives the sweep-midpoint energy the driver reported."""
    import json
    import subprocess
    import dmrgx_loader
    import bench_workload as W
    ham = ["-Lx", "4", "-Ly", "4", "-J1", "0.5", "-Jz1", "1", "-J2", "0.25", "-Jz2", "0.5", "-H_eps_tol", "1e-12", "-do_correlators", "0"]
    s1 = str(tmp_path) + "/s/"
    r = subprocess.run([EXE] + ham + ["-mwarmup", "24", "-msweeps", "32", "-scratch_dir", s1, "-data_dir", str(tmp_path) + "/d1/"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    # (kept-state counts inside degenerate +q / -q multiplets may differ between the two runs — SURVEY.md §7 — so the schedule
    # columns and the energies are compared, not the sector bookkeeping)
    r2 = subprocess.run([EXE, "-restart_dir", s1, "-msweeps", "256", "-H_eps_tol", "1e-12", "-do_correlators", "0", "-data_dir", str(tmp_path) + "/d2/"],
                        capture_output=True, text=True)
    assert r2.returncode == 0 and "Loading blocks from file" in r2.stdout, r2.stdout[-1500:] + r2.stderr[-1500:]
    r3 = subprocess.run([EXE] + ham + ["-mwarmup", "24", "-msweeps", "32,256", "-data_dir", str(tmp_path) + "/d3/"], capture_output=True, text=True)
    assert r3.returncode == 0
    t2 = json.load(open(str(tmp_path) + "/d2/DMRGSteps.json"))["table"]
    t3 = json.load(open(str(tmp_path) + "/d3/DMRGSteps.json"))["table"]
    assert len(t2) == 12
    for a, b in zip(t2, t3[-12:]):
        assert a[:8] == b[:8] and abs(a[-1] - b[-1]) <= 1e-8 * abs(b[-1]), (a, b)
    mid = [r for r in t2 if r[4] == r[5]]
    assert mid and abs(mid[-1][-1] - (-8.261232563030)) < 1e-7        # exact diagonalisation
    # InitializeFromDisk -> enlarge -> superblock -> ground state == the energy of that step in the driver's table
    P = dmrgx_loader.load_package()
    ctx = P.Context(0)
    W.CONFIGS["j1j2_4x4"] = dict(Lx=4, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5, bcx=0, bcy=1)
    wl = W.DiskWorkload(P, ctx, "j1j2_4x4", s1 + "Sweep_000000001/")
    e, psi, st = wl.shell.EPSSolve(tol=1e-12)
    last = json.load(open(str(tmp_path) + "/d1/DMRGSteps.json"))["table"][-1]
    assert wl.m == last[8] and wl.n == last[14] and abs(e - last[-1]) <= 1e-10 * abs(e)
    ctx.close()
