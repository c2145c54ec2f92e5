"""Dev tool: per-step wall time of the end-to-end path (pinned host batch -> H2D -> train_step -> loss.item())."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
import torch, bench, deepgate
from oracle import dg_oracle as O
w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4, layernorm=True)
model = deepgate.dg_ae_model_aig.Model(struct_encoder=enc, num_rounds=1, dim_hidden=64)
model.load_state_dict(O.synth_state_dict("aig", 2), strict=False)
tr = deepgate.Trainer(None, model, training_id="p", save_dir=tempfile.mkdtemp(), device=str(dev), distributed=False,
                      rc_prob_func_weight=[1.0, 4.0, 4.0])
model.train()
host = [bench.make_host_batch(w, 0, i).pin_memory() for i in range(4)]
def step(i):
    t0 = time.perf_counter()
    b = host[i % 4].copy_to(dev, non_blocking=True)
    t1 = time.perf_counter()
    st = tr.train_step(b)
    t2 = time.perf_counter()
    v = float(st["loss"].item())
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for i in range(5): step(i)
ts = [step(i) for i in range(40)]
import statistics
for k, name in enumerate(("h2d issue", "train_step host", "loss.item() wait")):
    xs = sorted(t[k] for t in ts)
    print("%-18s median %.2f  p90 %.2f  max %.2f ms" % (name, statistics.median(xs), xs[int(0.9 * len(xs))], xs[-1]))
tot = sorted(sum(t) for t in ts)
print("step total median %.2f p90 %.2f max %.2f ms; threads %d" % (statistics.median(tot), tot[int(0.9 * len(tot))], tot[-1], torch.get_num_threads()))
print("alloc retries", torch.cuda.memory_stats()["num_alloc_retries"], "cudaMalloc calls", torch.cuda.memory_stats()["num_device_alloc"], "reserved GB", torch.cuda.memory_reserved() / 1e9)
