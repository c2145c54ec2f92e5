"""Dev tool: per-phase clock64 trace of one backward step kernel (step k = 2)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from deepgate import _native as nat
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True).to(dev)
G = bench.make_host_batch(w, 0, 0).copy_to(dev, non_blocking=False)
feat = torch.nn.functional.one_hot(G.x[:, 1].to(torch.int64), num_classes=6).to(torch.float32)
lib = nat.lib()
lib.mgv_debug_set_trace.argtypes = [ctypes.c_void_p]
for i in range(2):
    s, t = enc(feat, feat, G.edge_index); (s.sum() + t.sum()).backward()
s, t = enc(feat, feat, G.edge_index)
tr = torch.zeros(2 * 74 * 16 * 16, dtype=torch.int64, device=dev)
lib.mgv_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
(s.sum() + t.sum()).backward()
torch.cuda.synchronize()
lib.mgv_debug_set_trace(ctypes.c_void_p(0))
t = tr.view(2 * 74, 16, 16).cpu().double()
names = ["gather", "step GEMM + gates", "LN backward", "GRU bwd + amax + DG store", "data-grad GEMMs + stores", "weight-grad GEMMs", "end barrier"]
valid = (t[:, :, 0] > 0) & (t[:, :, 7] > 0)
for i, n in enumerate(names):
    d = (t[:, :, i + 1] - t[:, :, i])[valid]
    print("%-28s mean %7.0f  med %7.0f  max %7.0f cycles" % (n, d.mean(), d.median(), d.max()))
d = (t[:, :, 7] - t[:, :, 0])[valid]
print("tile total mean %.0f cycles" % d.mean())
