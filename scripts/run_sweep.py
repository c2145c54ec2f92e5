"""Dev tool: the level sweep alone (forward + backward) on one synthetic batch, for ncu captures and event timing.
    python scripts/run_sweep.py [workload] [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from deepgate import ops
from deepgate.schedule import schedule_for_batch
from oracle import dg_oracle as O
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
ops.PROFILE = {}
for i in range(iters):
    hf = ops.level_sweep(hs, sch, w["rounds"], codes, mods)
    hf.sum().backward()
torch.cuda.synchronize()
for k, (n, ms) in ops.profile_summary().items():
    print("%s: %.3f ms per call, %.2f us per level (L=%d, N=%d, streams=%d)" % (k, ms / n, 1e3 * ms / n / max(sch.L - 1, 1), sch.L, sch.N, sch.streams))
