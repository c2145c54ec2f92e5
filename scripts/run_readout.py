"""Dev tool: the fused readout head alone (forward + backward) at a given N, for ncu captures.  python scripts/run_readout.py [N] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch
from deepgate import ops
from deepgate.arch.mlp import MLP
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65818
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
mlp = MLP(64, 32, 1, num_layer=3, p_drop=0.2, norm_layer="batchnorm", act_layer="relu").cuda().train()
x = torch.randn(N, 64, device="cuda", requires_grad=True)
t = torch.rand(N, 1, device="cuda")
ops.PROFILE = {}
for i in range(iters):
    pred, loss = ops.readout_head(x, t, mlp)
    loss.backward()
torch.cuda.synchronize()
for k, (n, ms) in ops.profile_summary().items():
    print("%s: %.1f us per call" % (k, 1e3 * ms / n))
