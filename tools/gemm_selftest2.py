"""Efficiency of ragged tile shapes: M = one tile row of height tm, N = 64 * 2000 (or transposed), K = 512."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmrgx_loader

P = dmrgx_loader.load_package()
ctx = P.Context(0)
L = P.lib()
def run(M, N, K, nseg=1, ak=1, bk=1):
    ms, err = C.c_double(), C.c_double()
    e = L.dmrgx_selftest_gemm(ctx.h, C.c_longlong(M), C.c_longlong(N), C.c_longlong(K), ak, bk, nseg, 5, C.byref(ms), C.byref(err))
    assert e == 0, L.dmrgx_last_error()
    return 2.0 * M * N * K * nseg / (ms.value * 1e-3) / 1e12, ms.value
for K in (512, 272):
    for tm in (64, 56, 48, 40, 32, 24, 16, 8):
        a, _ = run(tm, 64 * 4000, K)
        b, _ = run(64 * 4000, tm, K)
        print("K %4d tile %2dx64: %.2f TFLOP/s (%.2f of full-tile DMMA rate)   64x%2d: %.2f" % (K, tm, a, a / 34.4 / (((tm + 7) // 8 * 8) / 64.0), tm, b))
