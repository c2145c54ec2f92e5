"""Dev tool: per-phase clock64 trace of the last forward step kernel (mgv_debug_set_trace)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from deepgate import _native as nat
from deepgate.schedule import schedule_for_batch
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True).to(dev)
G = bench.make_host_batch(w, 0, 0).copy_to(dev, non_blocking=False)
sch = schedule_for_batch(G)
feat = torch.nn.functional.one_hot(G.x[:, 1].to(torch.int64), num_classes=6).to(torch.float32)
lib = nat.lib()
lib.mgv_debug_set_trace.argtypes = [ctypes.c_void_p]
for i in range(3):
    enc(feat, feat, G.edge_index)
tr = torch.zeros(2 * 74 * 16 * 16, dtype=torch.int64, device=dev)
lib.mgv_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
enc(feat, feat, G.edge_index)
torch.cuda.synchronize()
lib.mgv_debug_set_trace(ctypes.c_void_p(0))
t = tr.view(2 * 74, 16, 16).cpu().double()
names = {(0, 1): "G load issue", (1, 2): "G wait a_empty", (2, 3): "G h/x store (h loads land)", (3, 4): "G neighbour loop + agg store",
         (5, 6): "M wait acc_empty + a_full", (6, 7): "M issue", (8, 9): "E wait acc_full", (9, 10): "E gates", (10, 11): "E LN+store"}
valid = t[:, :, 0] > 0
print("tiles per CTA: min %d max %d" % (valid.sum(1).min(), valid.sum(1).max()))
for (a, b), n in names.items():
    d = (t[:, :, b] - t[:, :, a])[valid & (t[:, :, b] > 0) & (t[:, :, a] > 0)]
    print("%-32s mean %7.0f  med %7.0f  max %7.0f cycles" % (n, d.mean(), d.median(), d.max()))
# per tile totals for gather leader: loop top to next loop top
for cta in (0, 1, 36, 73, 74, 147):
    row = t[cta]; v = row[:, 0] > 0
    tops = row[v, 0]
    print("cta", cta, "tiles", int(v.sum()), "tile period", [(int(x)) for x in (tops[1:] - tops[:-1]).tolist()], "it0 G", int(row[0,4]-row[0,0]), "last E end - first G", int(row[v, 11].max() - row[0, 0]))
