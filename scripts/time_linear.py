"""Dev tool: csrc/linear_tc.cu against the library GEMM for the layer shapes of the path (CUDA events, 50 repetitions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch
from deepgate import ops
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for N in (65818, 6401024):
    for I, O in ((128, 64), (64, 128), (64, 64)):
        x = torch.randn(N, I, device="cuda"); W = torch.randn(O, I, device="cuda") * 0.1; b = torch.randn(O, device="cuda"); gy = torch.randn(N, O, device="cuda")
        tc_f = t(lambda: ops._linear_tc(x, W, b, O, I, False)); lib_f = t(lambda: torch.addmm(b, x, W.t()))
        tc_b = t(lambda: ops._linear_tc(gy, W, None, I, O, True)); lib_b = t(lambda: gy @ W)
        gb = N * (I + O) * 4 / 1e9
        print("N %8d %3d->%3d  forward tc %8.1f us (%.0f GB/s)  library %8.1f us | data gradient tc %8.1f us  library %8.1f us" % (N, I, O, tc_f, gb / tc_f * 1e6, lib_f, tc_b, lib_b))
