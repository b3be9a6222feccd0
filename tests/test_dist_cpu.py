"""world_size > 1 on CPU (gloo): the multi-GPU host logic of the product — KronSumShellSplitOwnership's successor (row
ranges cut at left-row boundaries), the sharded two-stage plan, the all-gather before an apply, all-reduced Lanczos
coefficients, eigen-blocks and rotated operators dealt to ranks and broadcast — checked against the oracle."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(world, config, m, mkeep):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "dmrg.x_b200", "csrc"), "plancheck"])
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), config, str(m), str(mkeep)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = []
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-3000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
        res.append(json.loads(line[7:]))
    return sorted(res, key=lambda d: d["rank"])


@pytest.mark.parametrize("world,config,m", [(2, "j1j2_12x6", 48), (3, "heis_8x4", 40)], ids=["world2-j1j2", "world3-heis"])
def test_sharded_path_matches_oracle(world, config, m):
    res = launch(world, config, m, (3 * m) // 4)
    n = res[0]["n"]
    cuts = res[0]["cuts"]
    assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts, cuts[1:]))
    assert sum(1 for a, b in zip(cuts, cuts[1:]) if b > a) == world  # every rank got work
    for r in res:
        assert r["cuts"] == cuts and r["range"] == [cuts[r["rank"]], cuts[r["rank"] + 1]]
        assert r["x_gathered"] and r["untouched"]
        assert r["matvec_err"] < 1e-12 and r["host_err"] < 1e-12
        assert r["converged"] and abs(r["e0"] - r["e_ref"]) <= 1e-10 * abs(r["e_ref"])
        assert abs(r["overlap"] - 1.0) < 1e-8 and abs(r["norm"] - 1.0) < 1e-12
        assert r["converged_from"] and abs(r["e0_from"] - r["e_ref"]) <= 1e-10 * abs(r["e_ref"]) and abs(r["overlap_from"] - 1.0) < 1e-8
        assert r["nmatvec_from"] <= r["nmatvec"]
        assert r["sectors_ok"] and abs(r["trunc_err"][0] - r["trunc_err"][1]) < 1e-12
        assert r["rot_H_err"] < 1e-11 and r["rot_Sp_err"] < 1e-11
        assert abs(r["expect"] - r["expect_ref"]) < 1e-12
    # all ranks took identical decisions
    assert len({(r["e0"], r["nmatvec"], r["e0_from"], r["nmatvec_from"]) for r in res}) == 1
