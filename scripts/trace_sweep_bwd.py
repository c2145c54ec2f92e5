"""Dev tool: accumulated per-phase cycles of the backward level sweep (mgv_debug_set_trace)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from deepgate import _native as nat, ops
from deepgate.schedule import schedule_for_batch
from oracle import dg_oracle as O
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True)
model = getattr(deepgate, "dg_ae_model_" + w["kind"]).Model(struct_encoder=enc, num_rounds=w["rounds"], dim_hidden=64)
model.load_state_dict(O.synth_state_dict(w["kind"], 2), strict=False)
model = model.to(dev)
G = bench.make_host_batch(w, 0, 0).copy_to(dev, non_blocking=False)
sch = schedule_for_batch(G)
hs = torch.randn(G.x.size(0), 64, device=dev, requires_grad=True)
codes = [c for c, _ in model.GATE_MODULES]
mods = [(getattr(model, "aggr_%s_func" % s), getattr(model, "update_%s_func" % s)) for _, s in model.GATE_MODULES]
lib = nat.lib()
lib.mgv_debug_set_trace.argtypes = [ctypes.c_void_p]
for i in range(2):
    hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods); hf.sum().backward()
hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods)
tr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.mgv_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
hf.sum().backward()
torch.cuda.synchronize()
lib.mgv_debug_set_trace(ctypes.c_void_p(0))
t = tr.view(148, 16).cpu().double()
names = ["P/A pull+gather", "B m GEMM", "C GRU+pointwise (incl. amax sync, DG store)", "D dm/dh GEMM", "E dxbar GEMM", "F attention bwd", "G wgrad",
         "pull-only nodes", "grid barrier wait", "loop overhead"]
tot = t[:, :10].sum(1)
print("levels", sch.L, "nodes", sch.N, "cycles/CTA mean %.0f" % tot.mean())
for i, n in enumerate(names):
    print("%-46s mean %10.0f cycles  %5.1f%%" % (n, t[:, i].mean(), 100 * t[:, i].mean() / tot.mean()))
cs = sch.code_count
print("code counts", cs)
import numpy as np
busy = (tot - t[:, 8]).numpy()
wait = t[:, 8].numpy()
order = np.argsort(busy)
print("busy cycles per CTA: min %.3g  p10 %.3g  median %.3g  p90 %.3g  max %.3g" % (busy.min(), np.percentile(busy, 10), np.median(busy), np.percentile(busy, 90), busy.max()))
for lo, hi in ((0, 37), (37, 74), (74, 111), (111, 148)):
    print("CTAs %3d-%3d: busy mean %.3g wait mean %.3g  P/A %.3g  F %.3g" % (lo, hi, busy[lo:hi].mean(), wait[lo:hi].mean(), t[lo:hi, 0].mean(), t[lo:hi, 5].mean()))
