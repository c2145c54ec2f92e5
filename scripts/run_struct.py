"""Dev tool: run the struct encoder (forward + backward) a few times on one cfg2-sized batch (for ncu / timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch
import bench
import deepgate
from deepgate import ops
from deepgate.schedule import schedule_for_batch

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True).to(dev)
G = bench.make_host_batch(w, 0, 0).copy_to(dev, non_blocking=False)
sch = schedule_for_batch(G)
feat = torch.nn.functional.one_hot(G.x[:, 1].to(torch.int64), num_classes=6).to(torch.float32)
ops.PROFILE = {}
for i in range(iters):
    s, t = enc(feat, feat, G.edge_index)
    (s.sum() + t.sum()).backward()
torch.cuda.synchronize()
for k, (n, ms) in ops.profile_summary().items():
    print(k, "calls", n, "ms/call", ms / n)
