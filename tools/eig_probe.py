"""The batched eigensolver on rho-like blocks of the sizes a 12x6 m=2048 (or --m) midpoint step has: the command ncu profiles."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmrgx_loader  # noqa: E402
import bench_workload as W  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = dmrgx_loader.load_package()
ctx = P.Context(0)
qn, sz = W.sector_sizes(m, 35)
enl = {}
for q, n in zip(qn, sz):
    for dq in (0.5, -0.5):
        enl[q + dq] = enl.get(q + dq, 0) + int(n)
sizes = [v for v in enl.values() if v > 0] * 2     # both sides
rng = np.random.default_rng(3)
mats = []
for n in sizes:
    k = max(1, min(n, int(0.9 * n)))
    X = rng.standard_normal((n, k)) * np.exp(-12.0 * np.arange(k) / k)[None, :]
    R = X @ X.T
    mats.append(R / np.trace(R))
n = np.array(sizes, np.int64)
for rep in range(reps):
    a = np.concatenate([M.ravel() for M in mats]); w = np.zeros(int(n.sum())); ms = C.c_double()
    l0 = P.launch_count()
    rc = P.lib().dmrgx_selftest_eig(ctx.h, C.c_longlong(len(mats)), n.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.byref(ms))
    assert rc == 0, P.lib().dmrgx_last_error()
    print("m=%d blocks=%s  solve %.2f ms  launches %d" % (m, sorted(sizes)[::-1][:8], ms.value, P.launch_count() - l0))
oa = ow = 0
worst = 0.0
for k, M in zip(n, mats):
    ref = np.linalg.eigvalsh(M)
    worst = max(worst, np.abs(w[ow:ow + k] - ref).max())
    oa += k * k; ow += k
print("max |lambda - numpy| = %.2e" % worst)
ctx.close()
