"""Dev tool: where does the host time of an end-to-end step go (run_batch / backward / optimizer), step by step?"""
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
tr = deepgate.Trainer(None, model, training_id="p", save_dir=tempfile.mkdtemp(), device=str(dev), distributed=False, rc_prob_func_weight=[1.0, 4.0, 4.0])
model.train()
host = [bench.make_host_batch(w, 0, i).pin_memory() for i in range(4)]
import gc
def step(i, log):
    t = [time.perf_counter()]
    b = host[i % 4].copy_to(dev, non_blocking=True); t.append(time.perf_counter())
    tr.optimizer.zero_grad(); t.append(time.perf_counter())
    st = tr.run_batch(b); t.append(time.perf_counter())
    loss = tr.total_loss(st); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    tr.grad_sync(); tr._guarded_step(); t.append(time.perf_counter())
    v = loss.item(); t.append(time.perf_counter())
    del b, st, loss; t.append(time.perf_counter())
    if log:
        print("step %2d: " % i + " ".join("%s %.2f" % (n, 1e3 * (t[k + 1] - t[k])) for k, n in enumerate(("h2d", "zero", "fwd", "loss", "bwd", "opt", "item", "free"))),
              " mallocs", torch.cuda.memory_stats(dev).get("num_device_alloc", 0))
for i in range(12): step(i, False)
torch.cuda.synchronize(); gc.collect(); gc.disable()
for rep in range(3):
    print("--- rep", rep)
    for i in range(12): step(i, True)
    time.sleep(0.5)
