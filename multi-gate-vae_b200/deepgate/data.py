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
_CIRCUIT_CSR = ("csr_in_src", "csr_in_deg", "csr_out_dst", "csr_out_slot", "csr_out_deg")     # per-circuit, merged by attach_host_schedule
CODE_SHIFT = 28           # out_pack = successor | code(successor) << 28   (include/mgv_b200.h)


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
        """A new batch object on ``device`` (the source, e.g. pinned host memory, is left untouched).  A packed batch
        (``pack()``) travels as ONE copy and its tensors are views of the device buffer."""
        out = OrderedData()
        packed = getattr(self, "_packed", None)
        if packed is not None and torch.device(device).type == "cuda":
            buf, table = packed
            dbuf = buf.to(device, non_blocking=non_blocking)
            out._dbuf = dbuf                            # the one device allocation behind every field
            for k, off, nbytes, dtype, shape in table:
                setattr(out, k, dbuf[off:off + nbytes].view(dtype).view(shape))
            for k, v in vars(self).items():
                if not k.startswith("_") and not torch.is_tensor(v):
                    setattr(out, k, v)
            return out
        for k, v in vars(self).items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pack(self, pin=True):
        """Lay all tensors of the (host) batch out in one contiguous, optionally pinned, byte buffer (256-byte aligned pieces) so
        that ``copy_to`` is a single host -> device copy instead of one per field (30 small copies cost 0.27 ms of host time per
        step); the fields become views of the buffer."""
        fields = [(k, v) for k, v in vars(self).items() if torch.is_tensor(v) and not k.startswith("_")]
        if any(v.is_cuda for _, v in fields):
            return self
        table, total = [], 0
        for k, v in fields:
            total = (total + 255) // 256 * 256
            nbytes = v.numel() * v.element_size()
            table.append((k, total, nbytes, v.dtype, tuple(v.shape)))
            total += nbytes
        buf = torch.empty(max(total, 1), dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            buf = buf.pin_memory()
        for (k, v), (_, off, nbytes, dtype, shape) in zip(fields, table):
            view = buf[off:off + nbytes].view(dtype).view(shape)
            view.copy_(v)
            setattr(self, k, view)
        self._packed = (buf, table)
        return self

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
        if key in _CIRCUIT_CSR:
            continue
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
    attach_host_schedule(out, circuits, starts)
    return out


SM_COUNT = 148             # B200
# Batches of at least this many circuits would run the level sweep in cluster mode.  Measured on B200 (cfg2: 64 AIG circuits of
# 500-1500 gates, 19 levels): cluster mode 0.155 / 0.543 ms forward / backward against 0.144 / 0.463 ms with two streams in grid
# mode -- without the grid barrier a level still costs one tile chain per circuit (~4 / 12 us), and a circuit's level is no
# longer spread over several CTAs.  So the default keeps two streams; ``MGV_SWEEP_STREAMS=<n>`` selects cluster mode.
CLUSTER_MIN_CIRCUITS = None


def choose_streams(batch, counts):
    """How many independent circuit sets ("streams") the level sweep gets (csrc/sweep_tc.cu):
      * many circuits (the reference's training batches): one set per thread-block cluster -- a cluster of one CTA per gate
        code walks its sets alone, levels separated by the hardware cluster barrier, no grid-wide synchronisation;
      * few circuits: two sets whose level chains overlap inside every CTA, grid barrier per set and level;
      * one circuit: a single stream.
    ``MGV_SWEEP_STREAMS`` overrides (development knob; 1, 2 or a cluster-mode count > 2)."""
    import os
    n = len(counts)
    env = os.environ.get("MGV_SWEEP_STREAMS")
    if env:
        return max(1, min(int(env), n))
    if n < 2:
        return 1
    if CLUSTER_MIN_CIRCUITS is None or n < CLUSTER_MIN_CIRCUITS:
        return 2
    codes = sum(1 for c in getattr(batch, "level_code_count", [1, 1]) if c > 0) or 1
    return max(3, min(n, SM_COUNT // codes))


def attach_streams(batch, counts, streams=None):
    """Cut the circuits of a batch into ``streams`` sets of near-equal size (largest first onto the lighter set) and record the
    set of every node in ``batch.sweep_stream`` (int32 [N]).  Circuits never exchange messages, so the level sweep may run the
    sets' level chains independently (mgv_b200.h, mgv_build_level_lists); the results do not depend on the cut.  Batches of one
    circuit keep a single stream."""
    streams = choose_streams(batch, counts) if streams is None else streams
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


def _stable_order(key, bound):
    """Stable argsort of small non-negative integer keys (radix sort in numpy for keys that fit 16 bits)."""
    import numpy as np
    k = key.astype(np.uint8 if bound <= 256 else (np.uint16 if bound <= 65536 else np.int64))
    return np.argsort(k, kind="stable").astype(np.int32)


def attach_host_schedule(batch, circuits, starts):
    """The batch's integer schedule, built on the HOST when the batch is collated (in the data-loader worker, where the
    reference builds ``forward_level``): in / out-edge CSR = the circuits' own CSRs (parser_func_others.attach_circuit_csr)
    concatenated with node / edge offsets, the (stream, level, code) node lists and the two degree orders.  The arrays travel
    with the batch (``sched_*``, int32) and ``schedule.schedule_for_batch`` wraps them -- the train step then launches no sort.
    Every array is bit-identical to what csrc/schedule.cu builds on the device (tests/test_gpu_parity.py)."""
    import numpy as np
    lvl = getattr(batch, "forward_level", None)
    gate = getattr(batch, "gate", None)
    if lvl is None or gate is None or not all(getattr(c, "csr_in_src", None) is not None for c in circuits):
        return batch
    n = int(batch.x.size(0))
    if lvl.numel() != n or gate.shape[0] != n:
        return batch
    code_raw = gate.reshape(-1).numpy().astype(np.int64)
    code = np.where((code_raw < 0) | (code_raw > 6), 6, code_raw)
    estarts = [0]
    for c in circuits:
        estarts.append(estarts[-1] + int(c.csr_in_src.numel()))
    in_deg = np.concatenate([c.csr_in_deg.numpy() for c in circuits]).astype(np.int64)
    out_deg = np.concatenate([c.csr_out_deg.numpy() for c in circuits]).astype(np.int64)
    in_ptr = np.zeros(n + 1, dtype=np.int64); np.cumsum(in_deg, out=in_ptr[1:])
    out_ptr = np.zeros(n + 1, dtype=np.int64); np.cumsum(out_deg, out=out_ptr[1:])
    in_src = np.concatenate([c.csr_in_src.numpy().astype(np.int64) + s0 for c, s0 in zip(circuits, starts)]) if estarts[-1] else np.zeros(0, np.int64)
    out_dst = np.concatenate([c.csr_out_dst.numpy().astype(np.int64) + s0 for c, s0 in zip(circuits, starts)]) if estarts[-1] else np.zeros(0, np.int64)
    out_slot = np.concatenate([c.csr_out_slot.numpy().astype(np.int64) + e0 for c, e0 in zip(circuits, estarts)]) if estarts[-1] else np.zeros(0, np.int64)
    out_pack = out_dst | (code[out_dst] << CODE_SHIFT)
    L = int(batch.num_levels)
    sos = getattr(batch, "sweep_stream", None)
    streams = int(getattr(batch, "sweep_streams", 1)) if sos is not None else 1
    lv = lvl.reshape(-1).numpy().astype(np.int64)
    key = ((sos.numpy().astype(np.int64) if streams > 1 else 0) * L + lv) * 8 + code
    nkeys = streams * L * 8
    order = _stable_order(key, nkeys)
    seg_ptr = np.zeros(nkeys + 1, dtype=np.int64); np.cumsum(np.bincount(key, minlength=nkeys), out=seg_ptr[1:])
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32)))
    batch.sched_in_ptr, batch.sched_in_src = i32(in_ptr), i32(in_src if in_src.size else np.zeros(1))
    batch.sched_out_ptr, batch.sched_out_pack, batch.sched_out_slot = i32(out_ptr), i32(out_pack if out_pack.size else np.zeros(1)), i32(out_slot if out_slot.size else np.zeros(1))
    batch.sched_order, batch.sched_seg_ptr = i32(order if n else np.zeros(1)), i32(seg_ptr)
    batch.sched_deg_order_in = i32(_stable_order(255 - np.minimum(in_deg, 255), 256) if n else np.zeros(1))
    batch.sched_deg_order_out = i32(_stable_order(255 - np.minimum(out_deg, 255), 256) if n else np.zeros(1))
    batch.sched_streams = streams
    return batch


class DataLoader(torch.utils.data.DataLoader):
    """``torch_geometric.loader.DataLoader`` stand-in used by ``Trainer`` (trainer.py:189-195)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=collate, **kw)


class CudaPrefetcher(object):
    """Iterates host batches (an iterable of ``OrderedData``, e.g. a ``DataLoader``) and yields them on ``device``, with the
    host -> device copy of batch i + 1 issued on a side stream while the caller works on batch i.  The reference moves every
    batch with a blocking ``batch.to(device)`` inside the step (trainer.py:201); here the copy (pinned memory, asynchronous)
    hides behind the previous step.  The yielded tensors are safe to use on the current stream (event wait +
    ``record_stream``)."""

    def __init__(self, batches, device, pin=True):
        self.batches, self.device, self.pin = batches, torch.device(device), pin
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def _upload(self, host):
        if self.stream is None:
            return host.to(self.device), None
        if getattr(host, "_packed", None) is None and self.pin:
            host = host.pack()                         # one pinned buffer, one copy
        with torch.cuda.stream(self.stream):
            dev = host.copy_to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._upload(next(it))          # batch i + 1 travels while the caller computes on batch i
            except StopIteration:
                nxt = None
            if ev is not None:
                main = torch.cuda.current_stream(self.device)
                main.wait_event(ev)
                dbuf = getattr(cur, "_dbuf", None)
                for v in ([dbuf] if dbuf is not None else [v for v in vars(cur).values() if torch.is_tensor(v)]):
                    v.record_stream(main)
            yield cur

    def __len__(self):
        return len(self.batches)


class DeferredScalars(object):
    """Reads device scalars back without stalling the step that produced them.  ``push(*tensors)`` copies the values (one
    element each) into a pinned host slot on the current stream -- asynchronous copies, no kernel -- and returns the list of floats
    pushed ``lag`` calls earlier, or ``None`` while fewer than ``lag`` pushes are in flight; it blocks only if that older copy has not
    finished, and by then a whole later step is queued behind it, so the device never drains.  ``drain()`` returns the lists still
    in flight, oldest first.  The reference reads ``loss.item()`` in the step itself (trainer.py:219-226), which empties the queue
    once per step: measured 0.5 ms of a 4.0 ms step at cfg2."""

    def __init__(self, device, lag=1, width=8):
        self.device, self.lag, self.width = torch.device(device), max(int(lag), 0), int(width)
        self.cuda = self.device.type == "cuda"
        self.slots = [torch.empty(self.width, dtype=torch.float64, pin_memory=self.cuda) for _ in range(self.lag + 1)]
        self.pending = []                      # (slot index, count, event), oldest first
        self.count = 0

    def _read(self, entry):
        i, n, ev = entry
        if ev is not None:
            ev.synchronize()
        return self.slots[i][:n].tolist()

    def push(self, *tensors):
        if len(tensors) > self.width:
            raise ValueError("DeferredScalars: %d values > width %d" % (len(tensors), self.width))
        out = self._read(self.pending.pop(0)) if self.lag > 0 and len(self.pending) == self.lag else None
        i = self.count % (self.lag + 1)
        self.count += 1
        slot = self.slots[i]
        for k, t in enumerate(tensors):
            slot[k:k + 1].copy_(t.detach().reshape(1), non_blocking=True)       # dtype conversion happens in the copy
        ev = None
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self.pending.append((i, len(tensors), ev))
        if self.lag == 0:
            return self._read(self.pending.pop(0))
        return out

    def drain(self):
        out = [self._read(e) for e in self.pending]
        self.pending = []
        return out

