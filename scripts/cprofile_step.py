"""Dev tool: cProfile of the host side of a training step (where does the CPU time go?)."""
import cProfile, os, pstats, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from oracle import dg_oracle as O
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True)
model = getattr(deepgate, "dg_ae_model_" + w["kind"]).Model(struct_encoder=enc, num_rounds=w["rounds"], dim_hidden=64)
model.load_state_dict(O.synth_state_dict(w["kind"], 2), strict=False)
tr = deepgate.Trainer(None, model, training_id="p", save_dir=tempfile.mkdtemp(), device=str(dev), distributed=False,
                      rc_prob_func_weight=[1.0, 4.0, 4.0])
model.train()
batches = [bench.make_host_batch(w, 0, i).copy_to(dev, non_blocking=False) for i in range(2)]
def step(i):
    b = batches[i % 2]; b._mgv_schedule = None; b.train_pos_edge_index = None
    tr.train_step(b)
for i in range(5): step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(20): step(i)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
