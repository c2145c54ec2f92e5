"""Host-built integer schedule (data.attach_host_schedule: per-circuit CSRs from the parser merged by offset at collate time)
against a direct construction from the collated ``edge_index`` -- the definition csrc/schedule.cu implements on the device
(stable by original edge id inside a node, as utils/dag_utils.py:91-105 ``subgraph`` concatenates a node's edges)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]


def test_merged_circuit_csr_equals_global_sort(monkeypatch):
    import deepgate
    from deepgate import synth
    monkeypatch.setenv("MGV_SWEEP_STREAMS", "9")       # cluster-mode cut: one circuit set per circuit
    b = deepgate.circuits_to_batch(synth.make_circuits("xmg", 9, (4, 12), (30, 300), cfg=71, window=25))
    n, E = b.x.size(0), b.edge_index.size(1)
    src, dst = b.edge_index[0], b.edge_index[1]
    by_dst = torch.sort(dst, stable=True).indices
    by_src = torch.sort(src, stable=True).indices
    assert torch.equal(b.sched_in_src.long(), src[by_dst])
    assert torch.equal(b.sched_in_ptr.long(), torch.cat([dst.new_zeros(1), torch.bincount(dst, minlength=n).cumsum(0)]))
    assert torch.equal(b.sched_out_ptr.long(), torch.cat([dst.new_zeros(1), torch.bincount(src, minlength=n).cumsum(0)]))
    code = b.gate.reshape(-1).long()
    out_dst = dst[by_src]
    assert torch.equal(b.sched_out_pack.long(), out_dst | (code[out_dst] << 28))
    pos_in = torch.empty(E, dtype=torch.long)
    pos_in[by_dst] = torch.arange(E)
    assert torch.equal(b.sched_out_slot.long(), pos_in[by_src])
    # every out-edge's slot points at the matching in-edge entry
    assert torch.equal(b.sched_in_src.long()[b.sched_out_slot.long()], src[by_src])
    L, S = b.num_levels, b.sched_streams
    assert S == 9
    key = (b.sweep_stream.long() * L + b.forward_level.long()) * 8 + code
    assert torch.equal(b.sched_order.long(), torch.sort(key, stable=True).indices)
    assert torch.equal(b.sched_seg_ptr.long(), torch.cat([key.new_zeros(1), torch.bincount(key, minlength=S * L * 8).cumsum(0)]))
    indeg = torch.bincount(dst, minlength=n)
    assert torch.equal(b.sched_deg_order_in.long(), torch.sort(255 - indeg.clamp(max=255), stable=True).indices)
    outdeg = torch.bincount(src, minlength=n)
    assert torch.equal(b.sched_deg_order_out.long(), torch.sort(255 - outdeg.clamp(max=255), stable=True).indices)


def test_streams_cut_whole_circuits_evenly():
    import deepgate
    from deepgate import synth
    b = deepgate.circuits_to_batch(synth.make_circuits("aig", 7, 8, (50, 400), cfg=72))
    s = b.sweep_stream.long()
    # a circuit lies in ONE stream; every edge stays inside a stream; the two sets are balanced (largest-first greedy)
    for g in range(b.num_graphs):
        assert len(set(s[b.ptr[g]:b.ptr[g + 1]].tolist())) == 1
    assert torch.equal(s[b.edge_index[0]], s[b.edge_index[1]])
    sizes = torch.bincount(s, minlength=2).tolist()
    biggest = int((b.ptr[1:] - b.ptr[:-1]).max())
    assert abs(sizes[0] - sizes[1]) <= biggest
    one = deepgate.circuits_to_batch(synth.make_circuits("aig", 1, 8, 50, cfg=73))
    assert getattr(one, "sweep_stream", None) is None and one.sched_streams == 1
    # a cluster-mode cut (more than two sets) of many circuits stays balanced
    import os
    os.environ["MGV_SWEEP_STREAMS"] = "37"
    try:
        many = deepgate.circuits_to_batch(synth.make_circuits("aig", 100, 8, (20, 60), cfg=74))
    finally:
        del os.environ["MGV_SWEEP_STREAMS"]
    assert many.sched_streams == 37 and int(many.sweep_stream.max()) == 36
    sizes = torch.bincount(many.sweep_stream.long())
    assert int(sizes.max()) - int(sizes.min()) <= int((many.ptr[1:] - many.ptr[:-1]).max())


def test_prefetcher_yields_every_batch_in_order():
    import deepgate
    from deepgate import synth
    batches = [deepgate.circuits_to_batch(synth.make_circuits("aig", 2, 4, 20 + 5 * i, cfg=80 + i)) for i in range(3)]
    got = list(deepgate.CudaPrefetcher(batches, "cpu"))
    assert len(got) == 3 and len(deepgate.CudaPrefetcher(batches, "cpu")) == 3
    for a, b in zip(got, batches):
        assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.sched_in_src, b.sched_in_src)
    assert list(deepgate.CudaPrefetcher([], "cpu")) == []


def test_packed_batch_is_one_buffer_with_the_same_fields():
    """``OrderedData.pack()``: every tensor becomes a view of ONE byte buffer (256-byte aligned pieces), values unchanged; a copy of
    the packed batch has the same fields and the plain-Python schedule metadata."""
    import deepgate
    from deepgate import synth
    b = deepgate.circuits_to_batch(synth.make_circuits("mig", 3, 8, 40, cfg=90))
    before = {k: getattr(b, k).clone() for k in b.keys() if torch.is_tensor(getattr(b, k))}
    meta = (b.num_levels, list(b.level_code_count), b.sched_streams, b.num_graphs)
    b.pack(pin=False)
    buf, table = b._packed
    base = buf.data_ptr()
    for k, off, nbytes, dtype, shape in table:
        v = getattr(b, k)
        assert v.data_ptr() == base + off and off % 256 == 0 and v.dtype == dtype and tuple(v.shape) == shape
        assert torch.equal(v, before[k]), k
    assert set(k for k, *_ in table) == set(before)
    c = b.copy_to("cpu")
    assert all(torch.equal(getattr(c, k), before[k]) for k in before)
    assert (c.num_levels, list(c.level_code_count), c.sched_streams, c.num_graphs) == meta


def test_deferred_scalars_return_every_value_once_in_order():
    """``DeferredScalars``: values come back ``lag`` pushes late, in order, each exactly once; ``drain`` returns the rest."""
    from deepgate import DeferredScalars
    for lag in (0, 1, 2, 3):
        r = DeferredScalars("cpu", lag=lag)
        got = []
        for i in range(7):
            out = r.push(torch.tensor(float(i)), torch.tensor(10 * i, dtype=torch.int32))
            assert (out is None) == (lag > 0 and i < lag)
            if out is not None:
                got.append(out)
        got += r.drain()
        assert got == [[float(i), float(10 * i)] for i in range(7)], (lag, got)
        assert r.drain() == []
    with pytest.raises(ValueError):
        DeferredScalars("cpu", width=1).push(torch.tensor(1.0), torch.tensor(2.0))


def test_weighted_loss_sum_matches_the_expression_and_its_gradients():
    """``Trainer.total_loss``'s one-function weighted sum: value and gradients of w0 a + w1 b + w2 c (+ wk kl)."""
    from deepgate.trainer import _WeightedSum
    w = [1.0, 4.0, 4.0, 0.25]
    xs = [torch.tensor(v, requires_grad=True) for v in (0.7, 0.11, 2.5, 13.0)]
    ys = [x.detach().clone().requires_grad_(True) for x in xs]
    a = _WeightedSum.apply(torch.tensor(w), *xs)
    b = sum(wi * y for wi, y in zip(w, ys))
    assert abs(float(a) - float(b)) < 1e-6
    (3.0 * a).backward()
    (3.0 * b).backward()
    assert all(abs(float(x.grad) - float(y.grad)) < 1e-6 for x, y in zip(xs, ys))


def test_edge_split_on_a_host_batch_keeps_every_edge():
    """``split_edges`` on a CPU batch (data preparation): a permutation of the columns of ``edge_index``."""
    import deepgate
    from deepgate import synth
    from deepgate.trainer import split_edges
    b = deepgate.circuits_to_batch(synth.make_circuits("aig", 2, 6, 50, cfg=3))
    split_edges(b)
    key = lambda t: torch.sort(t[0] * (1 << 20) + t[1]).values
    assert b.train_pos_edge_index.shape == b.edge_index.shape and torch.equal(key(b.train_pos_edge_index), key(b.edge_index))
