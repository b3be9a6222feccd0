"""Runs the sparse-sector H*psi (ExactChainWorkload) a few times: the command profiled by ncu for profiles/r2_spmm_kernel.md."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmrgx_loader  # noqa: E402
import bench_workload as W  # noqa: E402

nhalf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
P = dmrgx_loader.load_package()
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
ctx = P.Context(0, ts.cuda_stream)
sw = W.ExactChainWorkload(P, ctx, nhalf)
x = ctx.vec(sw.n, sw.random_state()); y = ctx.vec(sw.n)
for _ in range(3):
    sw.shell.MatMult(x, y)
scratch = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
ts_ = []
for _ in range(reps):
    scratch.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sw.shell.MatMult(x, y); b.record()
    torch.cuda.synchronize()
    ts_.append(a.elapsed_time(b))
st = sw.shell.stats()
print("D=%d alg_bytes=%d tiles=%d  ms cold: min %.4f median %.4f" % (sw.n, st["alg_bytes"], st["tiles_stage2"], min(ts_), sorted(ts_)[len(ts_) // 2]))
