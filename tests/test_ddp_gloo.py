"""Multi-process host logic on CPU (gloo, world size 2): the flat-buffer gradient all-reduce of
Trainer (mean of the per-rank gradients, one exchange per step) and the
per-rank sharding of circuits.  The CUDA kernels are not involved (they have no CPU path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import deepgate
        from deepgate import synth
        from deepgate.trainer import FlatGradAllReduce
        torch.manual_seed(0)
        # same parameters on every rank, rank-dependent gradients; one tensor without a gradient
        params = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11)),
                  torch.nn.Parameter(torch.randn(3, 3))]
        g = torch.Generator().manual_seed(100 + rank)
        params[0].grad = torch.randn(7, 5, generator=g)
        params[1].grad = torch.randn(11, generator=g)
        sync = FlatGradAllReduce(params)
        sync()
        out[rank] = [p.grad.clone() for p in params]
        # data sharding: DistributedSampler semantics over a list of circuits, collate on the host
        graphs = [deepgate.parse_pyg_mlpgate(c["x"], c["edge_index"], c["prob"], c["tt_sim"], c["tt_pair_index"])
                  for c in synth.make_circuits("mig4", 8, 8, 40, cfg=77, n_pairs=8)]
        sampler = torch.utils.data.distributed.DistributedSampler(graphs, num_replicas=world, rank=rank, shuffle=False)
        loader = deepgate.DataLoader(graphs, batch_size=2, sampler=sampler, drop_last=True)
        seen = []
        for b in loader:
            assert b.num_graphs == 2 and b.num_levels == int(b.forward_level.max()) + 1
            assert int(b.edge_index.max()) < b.x.size(0)            # index keys were shifted per circuit
            seen.append(int(b.x.size(0)))
        out[("n", rank)] = seen
        # replicas: every rank builds a differently initialised model (the reference's train.py sets no seed); the Trainer
        # broadcasts rank 0's parameters and buffers at construction, and after one synchronised step they still agree
        import tempfile
        from deepgate.trainer import Trainer
        torch.manual_seed(1000 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
        net[1].running_mean.add_(float(rank + 1))
        tr = Trainer(None, net, training_id="r%d" % rank, save_dir=tempfile.mkdtemp(), lr=1e-2, device="cpu", distributed=False)
        out[("sd0", rank)] = {k: v.clone() for k, v in tr.model.state_dict().items()}
        xg = torch.Generator().manual_seed(7 + rank)
        tr.optimizer.zero_grad()
        tr.model(torch.randn(16, 6, generator=xg)).square().mean().backward()
        tr.grad_sync()
        tr._guarded_step()
        out[("sd1", rank)] = {k: v.clone() for k, v in tr.model.state_dict().items() if "running" not in k and "num_batches" not in k}
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_is_mean_of_rank_grads_and_shards_are_disjoint():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ref = []
    for rank in range(world):
        g = torch.Generator().manual_seed(100 + rank)
        ref.append([torch.randn(7, 5, generator=g), torch.randn(11, generator=g)])
    mean0 = (ref[0][0] + ref[1][0]) / 2
    mean1 = (ref[0][1] + ref[1][1]) / 2
    for rank in range(world):
        got = out[rank]
        assert torch.allclose(got[0], mean0, atol=1e-7) and torch.allclose(got[1], mean1, atol=1e-7)
        assert torch.equal(got[2], torch.zeros(3, 3))               # a parameter without grad gets the (zero) mean
    assert len(out[("n", 0)]) == 2 and len(out[("n", 1)]) == 2      # 8 circuits -> 4 per rank -> 2 batches of 2
    for tag in ("sd0", "sd1"):
        a, b = out[(tag, 0)], out[(tag, 1)]
        assert a.keys() == b.keys() and len(a) >= 4
        for k in a:
            assert torch.equal(a[k], b[k]), (tag, k)
    assert float(out[("sd0", 1)]["1.running_mean"][0]) == 1.0        # rank 0's buffer (0 + 1), not rank 1's
