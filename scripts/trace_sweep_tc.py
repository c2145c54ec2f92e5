"""Dev tool: accumulated per-phase cycles of the tcgen05 level-sweep kernels (worker thread 0 of every CTA, clock64), plus
CUDA-event times of the sweep alone.  Needs the trace build of the library:
    MGV_OUT=/root/repo/build/_trace MGV_NVCC_EXTRA=-DMGV_SWEEP_TRACE bash multi-gate-vae_b200/csrc/build.sh
    MGV_B200_LIB=/root/repo/build/_trace/libmgv_b200.so python scripts/trace_sweep_tc.py [workload] [fwd|bwd]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
which = sys.argv[2] if len(sys.argv) > 2 else "bwd"
if which == "fwd":
    os.environ["MGV_TRACE_SWEEP_FWD"] = "1"
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
for i in range(3):
    hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods); hf.sum().backward()
# event timing without the trace
ops.PROFILE = {}
for i in range(5):
    hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods); hf.sum().backward()
torch.cuda.synchronize()
for k, (n, ms) in ops.profile_summary().items():
    print("%s: %.3f ms per call, %.2f us per level" % (k, ms / n, 1e3 * ms / n / max(sch.L - 1, 1)))
ops.PROFILE = None
tr = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
lib.mgv_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods)
hf.sum().backward()
torch.cuda.synchronize()
lib.mgv_debug_set_trace(ctypes.c_void_p(0))
t = tr.view(148, 32).cpu().double()
if which == "fwd":
    names = ["prefetch (+tile iter)", "wait level barrier", "gather+attention", "wait MMA", "epilogue", "end-of-tile sync", "grid arrive"]
    ntile_col = 8
else:
    names = ["prefetch + wait wgrad(prev)", "R gather+attention", "wait level barrier", "P pull", "wait recompute MMA", "W pointwise+planes",
             "wait dxbar MMA (+rescale)", "X transpose", "A attention bwd", "end-of-tile sync", "grid arrive", "tile loop overhead"]
    ntile_col = 12
tot = t[:, :len(names)].sum(1)
print("levels", sch.L, "nodes", sch.N, "clock cycles/CTA mean %.0f  (%.1f us per level at 1.9 GHz)" % (tot.mean(), tot.mean() / 1.9e3 / max(sch.L - 1, 1)))
print("tiles per CTA: mean %.1f max %d" % (t[:, ntile_col].mean(), int(t[:, ntile_col].max())))
for i, n in enumerate(names):
    print("%-32s mean %10.0f cycles  %5.1f%%   per level %7.0f" % (n, t[:, i].mean(), 100 * t[:, i].mean() / tot.mean(), t[:, i].mean() / max(sch.L - 1, 1)))
codes_of = t[:, 15].long()
for c in sorted(set(codes_of.tolist())):
    sel = codes_of == c
    wait_col = 1 if which == "fwd" else 2
    busy = tot[sel] - t[sel, wait_col]
    print("code %d: %3d CTAs  busy cycles per level mean %7.0f max %7.0f   tiles mean %.1f max %d" % (
        c, int(sel.sum()), busy.mean() / max(sch.L - 1, 1), busy.max() / max(sch.L - 1, 1), t[sel, ntile_col].mean(), int(t[sel, ntile_col].max())))
if which == "bwd":
    for i, n in ((16, "P: scalar loads issued"), (17, "P: row loads issued"), (18, "P: wait + reduce"), (19, "P: u / d u"), (20, "P: stores")):
        print("%-32s mean %10.0f cycles  %5.1f%%   per level %7.0f" % (n, t[:, i].mean(), 100 * t[:, i].mean() / tot.mean(), t[:, i].mean() / max(sch.L - 1, 1)))
print("code counts", sch.code_count, "streams", sch.streams)
