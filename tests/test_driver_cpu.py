"""CPU check of the DMRG-SquareLattice driver's HOST logic (schedule of SingleDMRGStep calls, option parsing, JSON
writers): the same driver source is linked against the test-only emulation of the device layer and its step table is
compared with the oracle's DMRG loop.  The real executable is checked on the GPU by tests/test_driver_gpu.py."""
import os
import subprocess

import pytest

import libswitch

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


def read_petsc_mat(path, int_bytes=4):
    """PETSc binary AIJ matrix (MatView on a binary viewer): big-endian classid 1211216, M, N, nz, row lengths, columns, values."""
    import numpy as np
    raw = open(path, "rb").read()
    it = np.dtype(">i4" if int_bytes == 4 else ">i8")
    hdr = np.frombuffer(raw, it, 4)
    assert hdr[0] == 1211216
    M, N, nz = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 4 * int_bytes
    rowlen = np.frombuffer(raw, it, M, off); off += M * int_bytes
    col = np.frombuffer(raw, it, nz, off); off += nz * int_bytes
    val = np.frombuffer(raw, ">f8", nz, off); off += nz * 8
    assert off == len(raw)
    A = np.zeros((M, N))
    r = np.repeat(np.arange(M), rowlen)
    A[r, col] = val
    return A


def test_checkpoint_format_and_restart(exe, orc, tmp_path):
    """-scratch_dir writes the reference's on-disk block format (BlockInfo.dat, QuantumNumbers.dat, PETSc binary .mat) after
    every loop; -restart_dir continues from it and reproduces the uninterrupted run."""
    import json
    import numpy as np
    ham = ["-Lx", "12", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-H_eps_tol", "1e-12", "-do_correlators", "0"]
    s1 = str(tmp_path) + "/scratch1/"
    r = subprocess.run([exe] + ham + ["-mwarmup", "16", "-msweeps", "24", "-scratch_dir", s1, "-data_dir", str(tmp_path) + "/d1/"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    sweep_dir = s1 + "Sweep_000000001/"
    info = dict(l.split() for l in open(sweep_dir + "Sys_000000005/BlockInfo.dat"))
    assert info["NumBytesPetscInt"] == "4" and info["NumBytesPetscScalar"] == "8" and info["PetscUseComplex"] == "0"
    assert info["SpinTypeKey"] == "102" and info["NumSites"] == "6"
    qn = [l.split() for l in open(sweep_dir + "Sys_000000005/QuantumNumbers.dat")]
    assert sum(int(a) for a, _ in qn) == int(info["NumStates"]) and len(qn) == int(info["NumSectors"])
    # the saved matrices against the oracle's block
    d = orc.DMRG(Lx=12, Ly=1, heisenberg=1.0, bcx=0, bcy=0, eps_tol=1e-12)
    d.warmup(16); d.sweep(24)
    Ho = d.block(5).get_op_dense(orc.OP_H)
    Hd = read_petsc_mat(sweep_dir + "Sys_000000005/H_000000000.mat")
    # (the kept multiplets at a degenerate cut may differ between the two codes, the low end of the block spectrum may not)
    assert Hd.shape == Ho.shape and abs(np.linalg.eigvalsh(Hd)[0] - np.linalg.eigvalsh(Ho)[0]) < 1e-9 and np.abs(Hd - Hd.T).max() < 1e-12
    Sz = read_petsc_mat(sweep_dir + "Sys_000000005/Sz_000000003.mat")
    assert np.abs(Sz - Sz.T).max() < 1e-12 and abs(np.abs(np.linalg.eigvalsh(Sz)).max() - 0.5) < 1e-9
    keys = dict(l.split() for l in open(sweep_dir + "Sweep.dat"))
    assert keys["LoopIdx"] == "1" and keys["num_sys_blocks"] == "11" and keys["sys_ninit"] == "6" and keys["GlobIdx"] == "12"
    # restart from the checkpoint, one more sweep at m = 32 == the uninterrupted 24,32 run
    r2 = subprocess.run([exe, "-restart_dir", s1, "-msweeps", "32", "-H_eps_tol", "1e-12", "-do_correlators", "0", "-data_dir", str(tmp_path) + "/d2/"],
                        capture_output=True, text=True)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]
    assert "Loading blocks from file" in r2.stdout
    r3 = subprocess.run([exe] + ham + ["-mwarmup", "16", "-msweeps", "24,32", "-data_dir", str(tmp_path) + "/d3/"], capture_output=True, text=True)
    assert r3.returncode == 0
    t2 = json.load(open(str(tmp_path) + "/d2/DMRGSteps.json"))["table"]
    t3 = json.load(open(str(tmp_path) + "/d3/DMRGSteps.json"))["table"]
    assert len(t2) == 8 and t2[0][0] == 12 and t2[0][2] == 2                      # GlobIdx / LoopIdx continue
    for a, b in zip(t2, t3[-8:]):
        assert a[:15] == b[:15] and abs(a[-1] - b[-1]) <= 1e-10 * abs(b[-1])
    assert abs(t2[-1][-1] - (-5.142090632841)) < 1e-9


@pytest.mark.parametrize("int_bytes", [4, 8], ids=["petscint32", "petscint64"])
def test_isolated_matvec_from_disk_saved_blocks(exe, tmp_path, int_bytes):
    """BASELINE configs[4]: blocks saved by the driver are read back through Block.InitializeFromDisk and the sweep-midpoint
    superblock built from them has the energy the driver reported for that step."""
    import json
    import dmrgx_loader
    import bench_workload as W
    args = ["-Lx", "4", "-Ly", "4", "-J1", "0.5", "-Jz1", "1", "-J2", "0.25", "-Jz2", "0.5", "-mwarmup", "24", "-msweeps", "32", "-H_eps_tol", "1e-12",
            "-do_correlators", "0", "-scratch_dir", str(tmp_path) + "/s/", "-data_dir", str(tmp_path) + "/d/", "-petsc_int_bytes", str(int_bytes)]
    r = subprocess.run([exe] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    info = dict(l.split() for l in open(str(tmp_path) + "/s/Sweep_000000001/Sys_000000006/BlockInfo.dat"))
    assert info["NumBytesPetscInt"] == str(int_bytes)
    P = dmrgx_loader.load_package()
    libswitch.use_library(P, os.path.join(ROOT, "tests", "plancheck", "libdmrgx_plancheck.so"))
    try:
        ctx = P.Context(0)
        W.CONFIGS["j1j2_4x4"] = dict(Lx=4, Ly=4, J1=0.5, Jz1=1.0, J2=0.25, Jz2=0.5, bcx=0, bcy=1)
        wl = W.DiskWorkload(P, ctx, "j1j2_4x4", str(tmp_path) + "/s/Sweep_000000001/")
        e, psi, st = wl.shell.EPSSolve(tol=1e-12)
        last = json.load(open(str(tmp_path) + "/d/DMRGSteps.json"))["table"][-1]
        assert wl.m == last[8] and wl.n == last[14]
        assert abs(e - last[-1]) <= 1e-10 * abs(e)
        ctx.close()
    finally:
        libswitch.use_library(P, None)


def test_spin_one_chain_matches_oracle(exe, orc, tmp_path):
    """-spin 1 (src/DMRGBlock.cpp:54-94, 1151-1156, 1210-1215): three-state sites, sector steps of one."""
    docs, out = dc.run_driver(exe, tmp_path, ["-Lx", 8, "-Ly", 1, "-heisenberg", 1, "-BCopen", "-spin", 1, "-do_correlators", 0], 27, [40])
    ref, _ = dc.compare_with_oracle(orc, docs, dict(Lx=8, Ly=1, heisenberg=1.0, bcx=0, bcy=0, spin_twice=2), 27, [40])
    # kept-state counts inside degenerate SU(2) multiplets may differ after the first tie; the converged energy may not
    assert abs(docs["DMRGSteps"]["table"][-1][-1] - ref[-1]["GSEnergy"]) <= 1e-9 * abs(ref[-1]["GSEnergy"])


def test_block_files_are_validated(exe, tmp_path):
    """InitializeFromDisk refuses what the reference refuses (src/DMRGBlock.cpp:256-276, 330-341): missing files, wrong scalar
    size, wrong sector count — and a .mat file that is not a PETSc binary matrix."""
    import shutil
    import dmrgx_loader
    args = ["-Lx", "8", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-mwarmup", "8", "-do_correlators", "0", "-scratch_dir", str(tmp_path) + "/s/",
            "-data_dir", str(tmp_path) + "/d/"]
    assert subprocess.run([exe] + args, capture_output=True, text=True).returncode == 0
    src = str(tmp_path) + "/s/Sweep_000000000/Sys_000000002/"
    P = dmrgx_loader.load_package()
    libswitch.use_library(P, os.path.join(ROOT, "tests", "plancheck", "libdmrgx_plancheck.so"))
    try:
        ctx = P.Context(0)
        good = P.Block.InitializeFromDisk(ctx, src)
        assert good.NumSites() == 3 and good.CheckOperatorBlocks() == 0

        def broken(name, edit):
            dst = str(tmp_path) + "/" + name + "/"
            shutil.copytree(src, dst)
            edit(dst)
            return dst
        d1 = broken("nofile", lambda d: os.remove(d + "Sp_000000001.mat"))
        with pytest.raises(Exception):
            P.Block.InitializeFromDisk(ctx, d1)
        d2 = broken("scalar", lambda d: open(d + "BlockInfo.dat", "w").write(open(src + "BlockInfo.dat").read().replace("NumBytesPetscScalar            8", "NumBytesPetscScalar            16")))
        with pytest.raises(P.DmrgxError):
            P.Block.InitializeFromDisk(ctx, d2)
        d3 = broken("classid", lambda d: open(d + "Sz_000000000.mat", "r+b").write(b"\\x00\\x00\\x00\\x01"))
        with pytest.raises(P.DmrgxError):
            P.Block.InitializeFromDisk(ctx, d3)
        # the C++ reader, through -restart_dir: a truncated QuantumNumbers.dat stops the run with an error
        bad = str(tmp_path) + "/s_bad/"
        shutil.copytree(str(tmp_path) + "/s/", bad)
        q = bad + "Sweep_000000000/Sys_000000001/QuantumNumbers.dat"
        first = open(q).read().splitlines()[0]
        open(q, "w").write(first + "\n")
        r = subprocess.run([exe, "-restart_dir", bad, "-msweeps", "8", "-data_dir", str(tmp_path) + "/d2/"], capture_output=True, text=True)
        assert r.returncode != 0 and "QuantumNumbers.dat" in r.stderr
        ctx.close()
    finally:
        libswitch.use_library(P, None)


def test_restart_takes_the_spin_type_from_the_block_files(exe, tmp_path):
    """src/DMRGBlock.cpp:280-315: restarting a spin-1 checkpoint WITHOUT -spin imposes the file's SpinTypeKey as -spin (the
    added sites must be spin-1 sites), and a contradicting -spin is an error."""
    import json
    ham = ["-Lx", "8", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-H_eps_tol", "1e-12", "-do_correlators", "0"]
    s = str(tmp_path) + "/s/"
    r = subprocess.run([exe] + ham + ["-spin", "1", "-mwarmup", "27", "-msweeps", "30", "-scratch_dir", s, "-data_dir", str(tmp_path) + "/d1/"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    info = dict(l.split() for l in open(s + "Sweep_000000001/Sys_000000002/BlockInfo.dat"))
    assert info["SpinTypeKey"] == "101"
    assert not os.path.exists(s + "Sweep_000000001/Sweep.dat.tmp") and os.path.exists(s + "Sweep_000000001/Sweep.dat")
    r2 = subprocess.run([exe, "-restart_dir", s, "-msweeps", "40", "-H_eps_tol", "1e-12", "-do_correlators", "0", "-data_dir", str(tmp_path) + "/d2/"],
                        capture_output=True, text=True)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]
    r3 = subprocess.run([exe] + ham + ["-spin", "1", "-mwarmup", "27", "-msweeps", "30,40", "-data_dir", str(tmp_path) + "/d3/"], capture_output=True, text=True)
    assert r3.returncode == 0
    t2 = json.load(open(str(tmp_path) + "/d2/DMRGSteps.json"))["table"]
    t3 = json.load(open(str(tmp_path) + "/d3/DMRGSteps.json"))["table"]
    assert len(t2) > 0 and t2[-1][14] == t3[-1][14]          # three-state sites were added: same superblock dimension
    assert abs(t2[-1][-1] - t3[-1][-1]) <= 1e-9 * abs(t3[-1][-1])
    r4 = subprocess.run([exe, "-restart_dir", s, "-spin", "1/2", "-msweeps", "40", "-do_correlators", "0", "-data_dir", str(tmp_path) + "/d4/"],
                        capture_output=True, text=True)
    assert r4.returncode != 0 and "do not match" in r4.stderr


def test_malformed_block_matrix_is_rejected(exe, tmp_path):
    """a .mat whose row lengths are negative but still add up to nz, or whose nz exceeds the file, must be refused — not read
    out of bounds (ReadMat / dmrgx_block_set_operator)."""
    import shutil
    import struct
    args = ["-Lx", "8", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-mwarmup", "8", "-do_correlators", "0", "-scratch_dir", str(tmp_path) + "/s/",
            "-data_dir", str(tmp_path) + "/d/"]
    assert subprocess.run([exe] + args, capture_output=True, text=True).returncode == 0
    f = str(tmp_path) + "/s/Sweep_000000000/Sys_000000001/Sz_000000000.mat"
    raw = bytearray(open(f, "rb").read())
    M, nz = struct.unpack(">i", raw[4:8])[0], struct.unpack(">i", raw[12:16])[0]
    lens = list(struct.unpack(">%di" % M, raw[16:16 + 4 * M]))
    assert M >= 2 and nz >= 1

    def restart_with(edit, name):
        bad = str(tmp_path) + "/" + name + "/"
        shutil.copytree(str(tmp_path) + "/s/", bad)
        g = bad + "Sweep_000000000/Sys_000000001/Sz_000000000.mat"
        b = bytearray(raw)
        edit(b)
        open(g, "wb").write(bytes(b))
        return subprocess.run([exe, "-restart_dir", bad, "-msweeps", "8", "-do_correlators", "0", "-data_dir", bad + "out/"], capture_output=True, text=True)

    def neg_rows(b):
        l2 = list(lens); l2[0] += 2; l2[1] -= 2     # {a+2, b-2, ...}: still sums to nz, second row may go negative
        if l2[1] >= 0:
            l2[0] += l2[1] + 1; l2[1] = -1
        b[16:16 + 4 * M] = struct.pack(">%di" % M, *l2)
    r = restart_with(neg_rows, "neg")
    assert r.returncode != 0 and ("invalid row length" in r.stderr or "row lengths" in r.stderr)
    r = restart_with(lambda b: b.__setitem__(slice(12, 16), struct.pack(">i", nz + 1000000)), "huge")
    assert r.returncode != 0 and "shorter than its header" in r.stderr


def test_wavefunction_prediction_option(exe, tmp_path):
    """-wavefunction_prediction 1 (extension; the reference starts every EPSSolve from a random vector): same energies and kept
    states at every step, most sweep steps start from the transformed previous ground state, fewer H*psi in all."""
    import json
    args = ["-Lx", 6, "-Ly", 4, "-J1", 0.5, "-Jz1", 1, "-J2", 0.25, "-Jz2", 0.5, "-do_correlators", 0]
    runs = []
    for flag in (0, 1):
        d = str(tmp_path) + "/p%d/" % flag
        r = subprocess.run([exe] + [str(a) for a in args] + ["-mwarmup", "16", "-msweeps", "24,32", "-data_dir", d, "-wavefunction_prediction", str(flag)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        runs.append((json.load(open(d + "DMRGRun.json")), json.load(open(d + "DMRGSteps.json"))))
    (r0, s0), (r1, s1) = runs
    assert r0["StepsWithPredictedStart"] == 0 and r1["StepsWithPredictedStart"] >= 30          # 40 sweep steps, 2 x 2 without a chain
    assert r1["NumMatVecs"] < 0.8 * r0["NumMatVecs"]
    h = s0["headers"]; ie = h.index("GSEnergy")
    for a, b in zip(s0["table"], s1["table"]):
        assert abs(a[ie] - b[ie]) < 1e-7 * abs(a[ie]), (a, b)
        assert a[:ie - 2] == b[:ie - 2] or a[:8] == b[:8]
