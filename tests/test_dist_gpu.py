"""-m gpu, needs >= 2 GPUs (skipped otherwise): the real NCCL path.  Two ranks of DMRG-SquareLattice.x, launched the
way bench.py is (torch.distributed.run, one process per GPU), must reproduce the oracle's step table, and the sharded
H*psi of bench.py's workload must agree with the single-GPU one."""
import os
import subprocess
import sys

import pytest

import driver_common as dc

pytestmark = pytest.mark.gpu
EXE = os.path.join(dc.ROOT, "dmrg.x_b200", "DMRG-SquareLattice.x")


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(ngpus() < 2, reason="needs two GPUs")
def test_two_rank_driver_matches_oracle(orc, tmp_path):
    import json
    out = str(tmp_path) + "/data/"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", EXE, "-Lx", "16", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-mwarmup", "24", "-msweeps", "48",
           "-H_eps_tol", "1e-12", "-data_dir", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    docs = {f: json.load(open(out + f + ".json")) for f in ("DMRGSteps", "Timings", "EntanglementSpectra", "DMRGRun", "Correlations")}
    ref, _ = dc.compare_with_oracle(orc, docs, dict(Lx=16, Ly=1, heisenberg=1.0, bcx=0, bcy=0), 24, [48])
    assert abs(docs["DMRGSteps"]["table"][-1][-1] - (-6.911737145575)) < 1e-8   # exact diagonalisation, L = 16


@pytest.mark.skipif(ngpus() < 2, reason="needs two GPUs")
def test_two_rank_driver_with_wavefunction_prediction(tmp_path):
    """-wavefunction_prediction 1 on two ranks (every rank transforms the whole vector, the eigen-solve takes its own rows):
    the exact energy of the 16-site chain, most sweep steps started from the prediction."""
    import json
    out = str(tmp_path) + "/data/"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", EXE, "-Lx", "16", "-Ly", "1", "-heisenberg", "1", "-BCopen", "-mwarmup", "24", "-msweeps", "48,48",
           "-H_eps_tol", "1e-12", "-do_correlators", "0", "-wavefunction_prediction", "1", "-data_dir", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    run = json.load(open(out + "DMRGRun.json")); steps = json.load(open(out + "DMRGSteps.json"))
    assert run["StepsWithPredictedStart"] >= 16
    assert abs(steps["table"][-1][-1] - (-6.911737145575)) < 1e-8
