"""One rank of the world_size>1 CPU test (gloo): the product's HOST logic for the multi-GPU path — row ownership, the
sharded plan, distributed Lanczos, distributed truncation / rotation — linked against the test-only emulation of the
device layer whose collectives are served by torch.distributed over gloo.  Launched by tests/test_dist_cpu.py."""
import ctypes as C
import json
import os
import sys

import numpy as np

import libswitch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dmrgx_loader
    import bench_workload as W
    from oracle import oracle as O
    P = dmrgx_loader.load_package()
    lib_path = os.path.join(ROOT, "tests", "plancheck", "libdmrgx_plancheck.so")
    libswitch.use_library(P, lib_path)
    L = P.lib()

    def np_view(ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))

    @C.CFUNCTYPE(None, C.c_void_p, C.c_longlong)
    def allreduce(ptr, n):
        t = torch.from_numpy(np_view(ptr, n))
        dist.all_reduce(t)

    @C.CFUNCTYPE(None, C.c_void_p, C.c_longlong, C.c_int)
    def bcast(ptr, n, root):
        t = torch.from_numpy(np_view(ptr, n))
        dist.broadcast(t, src=root)

    L.plancheck_set_collectives(allreduce, bcast)
    uid = P.dist_unique_id()
    ctx = P.Context(0, None, rank, world, uid)
    out = {"rank": rank}

    config, m, mkeep = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    wl = W.Workload(P, ctx, config, m=m)
    n = wl.n
    b, e, cuts = wl.shell.row_range()
    out["cuts"] = cuts.tolist(); out["range"] = [b, e]; out["n"] = n
    # --- oracle reference (every rank computes it; small)
    _, kb = W.oracle_side(O, wl)
    osh = O.Shell(kb, wl.terms)
    x = wl.random_state(3)
    y_ref = osh.apply(x)
    # --- sharded matvec: device buffers hold only the local rows on entry
    xin = np.zeros(n); xin[b:e] = x[b:e]
    dx = ctx.vec(n, xin); dy = ctx.vec(n, np.zeros(n))
    wl.shell.MatMult_sharded(dx, dy)
    y = dy.get()
    # the sector halo: whatever arrived is psi, whatever did not arrive is still zero (and was not needed: see matvec_err)
    got = dx.get()
    out["x_gathered"] = bool(np.all((got == x) | (got == 0.0)) and np.array_equal(got[b:e], x[b:e]))
    out["x_received_frac"] = float(np.count_nonzero(got) / max(1, np.count_nonzero(x)))
    out["matvec_err"] = float(np.abs(y[b:e] - y_ref[b:e]).max() / np.abs(y_ref).max()) if e > b else 0.0
    out["untouched"] = bool(np.all(y[:b] == 0) and np.all(y[e:] == 0))
    # --- host-buffer entry point: local rows in, local rows out
    yl = wl.shell.MatMult_host(x[b:e].copy())
    out["host_err"] = float(np.abs(yl - y_ref[b:e]).max() / np.abs(y_ref).max()) if e > b else 0.0
    # --- distributed Lanczos
    e0, psi, st = wl.shell.EPSSolve(tol=1e-12)
    e_ref, psi_ref, _, _ = osh.eigs(tol=1e-12)
    ph = psi.get()
    out["e0"] = e0; out["e_ref"] = e_ref; out["converged"] = st["converged"]; out["nmatvec"] = st["nmatvec"]
    out["overlap"] = float(abs(ph @ psi_ref)); out["norm"] = float(np.linalg.norm(ph))
    # --- the same solve started from a caller's vector (extension, dmrgx_eigs_smallest_from): a perturbed copy of the ground state,
    #     whole vector on every rank, each rank takes its own rows
    rng2 = np.random.default_rng(5)
    g = psi_ref + 1e-3 * rng2.standard_normal(n) / np.sqrt(n)
    e1, psi1, st1 = wl.shell.EPSSolve(tol=1e-12, initial=ctx.vec(n, 3.0 * g))
    out["e0_from"] = e1; out["nmatvec_from"] = st1["nmatvec"]; out["converged_from"] = st1["converged"]
    out["overlap_from"] = float(abs(psi1.get() @ psi_ref))
    # --- distributed truncation + rotation on the same vector as the oracle
    pd = ctx.vec(n, psi_ref)
    for mk in range(mkeep, mkeep + 8):
        obL = O.Truncation(kb, psi_ref, mk, True); obR = O.Truncation(kb, psi_ref, mk, False)
        if not (obL.tie or obR.tie):
            break
    btL, btR = P.GetTruncation(wl.kron, pd, mk)
    out["sectors_ok"] = bool(btL.sectors()[1].tolist() == obL.sectors()[1].tolist() and btR.sectors()[1].tolist() == obR.sectors()[1].tolist())
    out["trunc_err"] = [btL.TruncErr, obL.trunc_err]
    new = P.RotateOperators(wl.enl, btL)
    U = btL.RotMatT()
    enl_o, _ = W.oracle_side(O, wl)
    HL = enl_o.get_op_dense(O.OP_H)
    out["rot_H_err"] = float(np.abs(new.get_operator_dense(P.OpH) - U @ HL @ U.T).max())
    i = wl.used[0]
    Sp = enl_o.get_op_dense(O.OP_SP, i)
    out["rot_Sp_err"] = float(np.abs(new.get_operator_dense(P.OpSp, i) - U @ Sp @ U.T).max())
    # --- correlator through the sharded single-term shell
    h1 = wl.kron.KronConstruct(P.OpSz, wl.used[0], P.OpSz, wl.used[0])
    out["expect"] = h1.expect(pd)
    o1 = O.Shell(kb, single=(O.OP_SZ, wl.used[0], O.OP_SZ, wl.used[0]))
    out["expect_ref"] = float(psi_ref @ o1.apply(psi_ref))
    print("RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
