"""Runs the chain engine on plain GEMMs (through dmrgx_selftest_gemm) to separate kernel efficiency from workload structure."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmrgx_loader

P = dmrgx_loader.load_package()
ctx = P.Context(0)
L = P.lib()
for (M, N, K, ak, bk, nseg) in [(4096, 4736, 2048, 1, 1, 1), (4096, 4736, 2048, 1, 0, 1), (4096, 4736, 2048, 0, 0, 1), (4096, 4736, 2048, 0, 1, 1),
                                (4096, 4736, 512, 1, 1, 4), (4096, 4736, 256, 1, 1, 8), (4096, 4736, 512, 1, 1, 1), (4096, 4736, 256, 1, 1, 1),
                                (4096 + 48, 4736 + 48, 2048, 1, 1, 1), (1024, 1024, 512, 1, 1, 1), (2048, 2048, 512, 1, 1, 1)]:
    ms, err = C.c_double(), C.c_double()
    e = L.dmrgx_selftest_gemm(ctx.h, C.c_longlong(M), C.c_longlong(N), C.c_longlong(K), ak, bk, nseg, 5, C.byref(ms), C.byref(err))
    assert e == 0, L.dmrgx_last_error()
    print("M %5d N %5d K %5d x%d  A %s B %s   %.3f ms  %.2f TFLOP/s  err %.1e" % (M, N, K, nseg, "mk" if ak else "km", "nk" if bk else "kn", ms.value,
                                                                        2.0 * M * N * K * nseg / (ms.value * 1e-3) / 1e12, err.value))
