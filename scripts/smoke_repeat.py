"""Dev tool: the smoke configuration (MIG, 4 x 216 nodes) run repeatedly in ONE process, with the error of every stage
against the CPU oracle (fp64) printed per repetition: struct encoder (s, t), hs_linear alone, level sweep alone (oracle fed
with the device's hs), and the end-to-end hs / hf / worst gradient.  Written to explain the 10x swing between two smoke runs of
round 1 (1.4e-5 plain, 1.4e-6 under ncu); see DESIGN.md section 3.

    python scripts/smoke_repeat.py [reps]           (also try CUDA_LAUNCH_BLOCKING=1)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import deepgate  # noqa: E402
from deepgate import synth  # noqa: E402
from oracle import dg_oracle as O  # noqa: E402
from util import build_model, oracle_inputs, oracle_train_grads, rel  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
circuits = synth.make_circuits("mig4", 4, 16, 200, cfg=1, n_pairs=32)
G = deepgate.circuits_to_batch(circuits, "cuda:0")
sd = O.synth_state_dict("mig", 11)
gen = torch.Generator().manual_seed(0)
E, n = G.edge_index.size(1), G.x.size(0)
pos, neg = G.edge_index.cpu()[:, torch.randperm(E, generator=gen)], torch.randint(0, n, (2, E), generator=gen)
inputs = oracle_inputs(G, pos, neg)
# fp64 oracle: the truth every stage is compared with
sd64 = {k: v.double() if v.is_floating_point() else v for k, v in sd.items()}
total64, parts64, grads64 = oracle_train_grads("mig", sd, inputs, (1.0, 4.0, 4.0), 1, dtype=torch.float64)
total32, parts32, grads32 = oracle_train_grads("mig", sd, inputs, (1.0, 4.0, 4.0), 1)
code = inputs["code"]
s64, t64 = O.struct_encoder(sd64, "mig_struct_encoder", code, inputs["edge_index"], 4, 4, True)
print("fp32 oracle vs fp64 oracle: hs %.2e hf %.2e" % (rel(parts32["hs"], parts64["hs"]), rel(parts32["hf"], parts64["hf"])))
rows = []
for rep in range(reps):
    model = build_model("mig", sd, 1, device="cuda:0")
    stash = {}
    hook = model.mig_struct_encoder.register_forward_hook(lambda m, i, o: stash.update(s=o[0].detach(), t=o[1].detach()))
    G._mgv_schedule = None
    hs, hf = model(G)               # in repetition 0 this is the FIRST device work of the process (as in smoke())
    hook.remove()
    s, t = stash["s"], stash["t"]
    hs_dev = hs.detach()
    hs_lin64 = torch.nn.functional.linear(torch.cat([s, t], -1).detach().cpu().double(), sd64["hs_linear.weight"], sd64["hs_linear.bias"])
    _, hf_given_hs = O.model_forward(sd64, "mig", code, inputs["edge_index"], inputs["forward_level"], 1, literal_subgraph=False,
                                     hs_override=hs.detach().cpu().double())
    rec, _, _ = model.recon_loss(hs, pos.cuda(), neg.cuda())
    prb = torch.nn.L1Loss()(model.pred_prob(hf), G.prob)
    _, _, _, fnc = deepgate.ops.vae_func_loss(hf=hf, tt_pair_index=G.tt_pair_index, tt_sim=G.tt_sim)
    (rec + 4 * prb + 4 * fnc).backward()
    worst = 0.0
    for k, p in model.named_parameters():
        ref = grads64.get(k)
        if ref is None or any(tag in k for tag in (".msg_q.", "msg_k.bias", "attn_lin.bias", "attn_lin.weight", "msg_k.weight")):
            continue
        worst = max(worst, rel(p.grad, ref))
    row = (rel(s, s64), rel(t, t64), rel(hs_dev, hs_lin64), rel(hf, hf_given_hs), rel(hs, parts64["hs"]), rel(hf, parts64["hf"]), worst)
    rows.append(row)
    print("rep %2d: s %.2e t %.2e | hs_linear alone %.2e | sweep alone %.2e | hs %.2e hf %.2e grad %.2e" % ((rep,) + row))
cols = list(zip(*rows))
names = ("s", "t", "hs_linear", "sweep", "hs", "hf", "grad")
print("max over %d reps: " % reps + "  ".join("%s %.2e" % (nme, max(c)) for nme, c in zip(names, cols)))
print("min over %d reps: " % reps + "  ".join("%s %.2e" % (nme, min(c)) for nme, c in zip(names, cols)))
