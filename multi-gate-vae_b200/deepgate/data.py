"""Batch container for circuits: the data boundary of ``Model.forward(G)``.

Mirrors what the reference gets from PyG's ``Batch`` of ``OrderedData``
(parser_func_others.py:10-40): a batch is the disjoint union of circuits; keys
whose name contains ``index`` are shifted by the running node count,
``edge_index`` / ``tt_pair_index`` concatenate along dim 1, everything else
along dim 0 un-shifted (so ``forward_level`` is shared across circuits).
PyTorch-Geometric is not a dependency of this package.
"""
import torch

_CAT_LAST = ("edge_index", "tt_pair_index", "rc_pair_index")


class OrderedData(object):
    """Attribute bag with the reference's batching rules (parser_func_others.py:28-40)."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    # mapping-style access, as PyG's Data offers (trainer.py:155-162 uses batch['prob'])
    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return getattr(self, key, None) is not None

    def keys(self):
        return [k for k, v in vars(self).items() if v is not None and not k.startswith("_")]

    @property
    def num_nodes(self):
        return int(self.x.size(0))

    def __inc__(self, key, value=None):
        return self.num_nodes if ("index" in key or "face" in key) else 0

    def __cat_dim__(self, key, value=None):
        return 1 if key in _CAT_LAST else 0

    def to(self, device, non_blocking=False):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def copy_to(self, device, non_blocking=True):
        """A new batch object on ``device`` (the source, e.g. pinned host memory, is left untouched)."""
        out = OrderedData()
        for k, v in vars(self).items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in vars(self).values() if torch.is_tensor(v))

    def pin_memory(self):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.pin_memory())
        return self


def collate(circuits):
    """Disjoint union of ``OrderedData`` circuits (adds ``batch`` and ``ptr``)."""
    first = circuits[0]
    out = OrderedData()
    counts = [c.num_nodes for c in circuits]
    starts = [0]
    for n in counts:
        starts.append(starts[-1] + n)
    for key in first.keys():
        vals = [c[key] for c in circuits]
        if not torch.is_tensor(vals[0]):
            out[key] = vals[0]
            continue
        if first.__inc__(key):
            vals = [v + s for v, s in zip(vals, starts)]
        out[key] = torch.cat(vals, dim=first.__cat_dim__(key))
    out.batch = torch.repeat_interleave(torch.arange(len(circuits)), torch.tensor(counts))
    out.ptr = torch.tensor(starts, dtype=torch.long)
    out.num_graphs = len(circuits)
    attach_schedule_meta(out)
    attach_streams(out, counts)
    return out


SWEEP_STREAMS = 2          # independent circuit sets the level sweep runs concurrently (csrc/sweep_tc.cu)


def attach_streams(batch, counts, streams=SWEEP_STREAMS):
    """Cut the circuits of a batch into ``streams`` sets of near-equal size (largest first onto the lighter set) and record the
    set of every node in ``batch.sweep_stream`` (int32 [N]).  Circuits never exchange messages, so the level sweep may run the
    sets' level chains concurrently, each with its own barrier (mgv_b200.h, mgv_build_level_lists); the results do not depend
    on the cut.  Batches of one circuit keep a single stream."""
    if len(counts) < 2 or streams < 2:
        return batch
    load = [0] * streams
    table = [0] * len(counts)
    for i in sorted(range(len(counts)), key=lambda i: -counts[i]):
        k = min(range(streams), key=lambda k: load[k])
        table[i] = k
        load[k] += counts[i]
    batch.sweep_stream = torch.repeat_interleave(torch.tensor(table, dtype=torch.int32), torch.tensor(counts))
    batch.sweep_streams = streams
    return batch


def attach_schedule_meta(batch):
    """Host-side schedule metadata, computed where the reference computes ``forward_level`` (at data-preparation
    time, parser_func_others.py:63): the number of levels and the nodes per gate code at level >= 1.  With these two
    plain-Python attributes on the batch, ``Model.forward`` builds its device schedule without a host sync (the
    reference syncs ~2N + 5L times per forward)."""
    lvl = getattr(batch, "forward_level", None)
    gate = getattr(batch, "gate", None)
    if lvl is None or gate is None or lvl.is_cuda or lvl.numel() == 0 or lvl.numel() != gate.shape[0]:
        return batch
    lvl = lvl.reshape(-1).to(torch.int64)
    code = gate.reshape(-1).to(torch.int64).clamp(0, 6)
    code = torch.where((gate.reshape(-1) < 0) | (gate.reshape(-1) > 6), torch.full_like(code, 6), code)
    batch.num_levels = int(lvl.max()) + 1
    batch.level_code_count = torch.bincount(code[lvl >= 1], minlength=8)[:8].tolist()
    return batch


class DataLoader(torch.utils.data.DataLoader):
    """``torch_geometric.loader.DataLoader`` stand-in used by ``Trainer`` (trainer.py:189-195)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=collate, **kw)
