"""Dev tool: one small train step (struct encoder + level sweep + losses, forward and backward) for compute-sanitizer runs.
    compute-sanitizer --tool racecheck python scripts/run_small.py [kind] [circuits] [gates]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, deepgate
from deepgate import synth
from oracle import dg_oracle as O
kind = sys.argv[1] if len(sys.argv) > 1 else "mig"
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ng = int(sys.argv[3]) if len(sys.argv) > 3 else 300
G = deepgate.circuits_to_batch(synth.make_circuits(kind, nc, 8, ng, cfg=9, window=24), "cuda")
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=2, t_rounds=2, layernorm=True)
model = getattr(deepgate, "dg_ae_model_" + kind).Model(struct_encoder=enc, num_rounds=1, dim_hidden=64)
model.load_state_dict(O.synth_state_dict(kind, 2), strict=False)
tr = deepgate.Trainer(None, model, training_id="s", save_dir=tempfile.mkdtemp(), device="cuda:0", distributed=False, rc_prob_func_weight=[1.0, 4.0, 4.0])
model.train()
for i in range(2):
    G._mgv_schedule = None
    G.train_pos_edge_index = None
    st = tr.train_step(G)
torch.cuda.synchronize()
print("loss", float(st["loss"]), "nodes", G.x.size(0), "levels", G.num_levels)
