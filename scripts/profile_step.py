"""Where does a training step spend host and device time?  (torch.profiler, dev tool)"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch
import bench
import deepgate
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
for i in range(3): step(i)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(10): step(i)
torch.cuda.synchronize()
print("ms/step wall", (time.perf_counter() - t0) * 100)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
