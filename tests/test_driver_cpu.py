"""CPU check of the DMRG-SquareLattice driver's HOST logic (schedule of SingleDMRGStep calls, option parsing, JSON
writers): the same driver source is linked against the test-only emulation of the device layer and its step table is
compared with the oracle's DMRG loop.  The real executable is checked on the GPU by tests/test_driver_gpu.py."""
import os
import subprocess

import pytest

import driver_common as dc

ROOT = dc.ROOT
EXE = os.path.join(ROOT, "tests", "plancheck", "DMRG-SquareLattice.plancheck.x")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "dmrg.x_b200", "csrc"), "plancheck"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "dmrg.x_b200", "driver"), "plancheck"])
    return EXE


def test_chain12_matches_oracle_and_exact_energy(exe, orc, tmp_path):
    docs, out = dc.run_driver(exe, tmp_path, ["-Lx", 12, "-Ly", 1, "-heisenberg", 1, "-BCopen"], 16, [24, 32])
    ref, _ = dc.compare_with_oracle(orc, docs, dict(Lx=12, Ly=1, heisenberg=1.0, bcx=0, bcy=0), 16, [24, 32])
    assert abs(ref[-1]["GSEnergy"] - (-5.142090632841)) < 1e-9           # SURVEY.md §8c, L = 12 open chain
    assert abs(docs["DMRGSteps"]["table"][-1][-1] - (-5.142090632841)) < 1e-9
    assert docs["DMRGRun"]["Sweeps"]["MStates"] == [24, 32]
    assert len(docs["Timings"]["table"]) == len(docs["DMRGSteps"]["table"]) == len(docs["EntanglementSpectra"])
    vals = dc.check_correlations(docs, 12)
    # Heisenberg point: <S+S-> = 2 <SzSz> on every bond
    assert abs(vals["NearestNeighborSpSm( 5 6 )"] - 2 * vals["NearestNeighborSzSz( 5 6 )"]) < 1e-6


def test_j1j2_cylinder_4x4_matches_oracle(exe, orc, tmp_path):
    args = ["-Lx", 4, "-Ly", 4, "-J1", 0.5, "-Jz1", 1, "-J2", 0.25, "-Jz2", 0.5]
    docs, out = dc.run_driver(exe, tmp_path, args, 24, [40])
    ref, _ = dc.compare_with_oracle(orc, docs, dict(Lx=4, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5), 24, [40])
    assert ref[-1]["GSEnergy"] >= -8.261232563030 - 1e-9                 # variational bound, SURVEY.md §8c
    dc.check_correlations(docs, 16)


def test_sweep_modes_and_errors(exe, tmp_path):
    base = [exe, "-Lx", "8", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-data_dir", str(tmp_path) + "/d/"]
    r = subprocess.run(base + ["-mwarmup", "8", "-nsweeps", "2"], capture_output=True, text=True)
    assert r.returncode == 0 and "SWEEP_MODE_NSWEEPS" in r.stdout and r.stdout.count("SWEEP MStates=8") == 2
    r = subprocess.run(base + ["-mwarmup", "8", "-msweeps", "8,12", "-maxnsweeps", "2,3"], capture_output=True, text=True)
    assert r.returncode == 0 and "SWEEP_MODE_TOLERANCE_TEST" in r.stdout and "BREAK" in r.stdout
    r = subprocess.run(base + ["-mwarmup", "8", "-msweeps", "8,12", "-nsweeps", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot both be specified" in r.stderr
    r = subprocess.run(base + ["-mwarmup", "8", "-msweeps", "8,12", "-maxnsweeps", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "same number of items" in r.stderr
    r = subprocess.run([exe, "-Lx", "3", "-Ly", "1", "-mwarmup", "8", "-data_dir", str(tmp_path) + "/e/"], capture_output=True, text=True)
    assert r.returncode != 0 and "must be even" in r.stderr
