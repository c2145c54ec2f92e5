"""Device-built level schedule of a circuit batch.

Replaces, on the GPU and once per batch, what the reference recomputes on the host for
every level, gate type and node: ``layer_mask & type_mask`` node selection
(dg_ae_model_mig.py:86-89) and ``subgraph`` (utils/dag_utils.py:91-105), plus ``top_sort``
(utils/dag_utils.py:10-37) when a batch carries no ``forward_level``.

Holds int32 device arrays:
  in_ptr/in_src      predecessors by node id, ascending original edge id
  out_ptr/out_pack/out_slot  successors (out_pack = dst | code(dst) << 28)
  level              ASAP level per node
  order/seg_ptr      node ids sorted by (stream, level, code), segment boundaries [streams*L*8+1]
                     (streams = independent circuit sets of the batch, data.attach_streams; 1 unless asked for)
"""
import ctypes
import os

import torch

from . import _native as nat


_ERR = {}          # device -> int32 word: deferred validation flags of the asynchronous schedule builds


def error_word(device):
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _ERR:
        _ERR[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return _ERR[key]


def check_deferred_errors():
    """Synchronise and raise if an asynchronous schedule build / fused loss saw invalid input since the last check
    (bit 0: node id outside [0, N); bit 1: level outside [0, L); bit 2: loss edge end outside [0, N))."""
    for key, word in _ERR.items():
        v = int(word.item())
        if v:
            word.zero_()
            raise RuntimeError("mgv_b200: invalid graph input on %s:%s (flags 0x%x: 1 = node id out of range, "
                               "2 = level out of range, 4 = loss edge out of range)" % (key[0], key[1], v))


class GraphCSR(object):
    """In/out-edge CSR of ``edge_index`` ([2, E] int64, row 0 = source).  ``validate=True`` checks node ids
    synchronously; ``False`` defers the check to ``check_deferred_errors()`` and never syncs."""

    @classmethod
    def from_host_schedule(cls, G):
        """Wrap the integer schedule the batch carries (data.attach_host_schedule: CSR, level lists and degree orders built
        on the host at collate time) -- no sort runs on the device; only the tile descriptors / costs and the sweep's row
        descriptors are derived here (7 small launches)."""
        self = cls.__new__(cls)
        ei = G.edge_index
        dev = ei.device
        self.device, self.edge_index = dev, ei
        self.N, self.E = int(G.gate.shape[0]), int(ei.size(1))
        self.code = nat.require_cuda(G.gate.reshape(-1).to(torch.int32).contiguous(), "code", torch.int32)
        for k in ("in_ptr", "in_src", "out_ptr", "out_pack", "out_slot", "order", "seg_ptr", "deg_order_in", "deg_order_out"):
            setattr(self, k, nat.require_cuda(getattr(G, "sched_" + k), "sched_" + k, torch.int32))
        self.level = G.forward_level
        self.L = max(int(G.num_levels), 1)
        self.streams = int(G.sched_streams)
        self.code_count = [int(c) for c in G.level_code_count]
        if self.in_ptr.numel() != self.N + 1 or self.order.numel() != max(self.N, 1) or \
                self.seg_ptr.numel() != self.streams * self.L * nat.NCODE + 1:
            raise RuntimeError("mgv_b200: the batch's host-built schedule does not match its nodes / levels")
        i32 = dict(dtype=torch.int32, device=dev)
        ntiles = (self.N + nat.TILE_ROWS - 1) // nat.TILE_ROWS
        self.tile_cost_in = torch.empty(ntiles + 1, **i32)
        self.tile_cost_out = torch.empty(ntiles + 1, **i32)
        self.gdesc_in = torch.empty(max(self.N, 1), 4, **i32)
        self.gdesc_out = torch.empty(max(self.N, 1), 4, **i32)
        self.sweep_desc = torch.empty(max(self.N, 1), 8, **i32)
        lib = nat.lib()
        with nat.on_device(dev):
            nb = lib.mgv_degree_order_workspace_bytes(self.N)
            ws = nat.workspace(nb, dev)
            st = nat.stream_of(dev)
            for p_, i_, o_, g_, c_ in ((self.in_ptr, self.in_src, self.deg_order_in, self.gdesc_in, self.tile_cost_in),
                                       (self.out_ptr, self.out_pack, self.deg_order_out, self.gdesc_out, self.tile_cost_out)):
                nat.check(lib.mgv_build_degree_tiles(nat.ptr(p_), nat.ptr(i_), nat.ptr(o_), self.N, nat.ptr(g_), nat.ptr(c_),
                                                     nat.ptr(ws), nb, st), "mgv_build_degree_tiles")
            nat.check(lib.mgv_build_sweep_desc(nat.ptr(self.order), nat.ptr(self.in_ptr), nat.ptr(self.in_src), nat.ptr(self.out_ptr),
                                               self.N, nat.ptr(self.sweep_desc), st), "mgv_build_sweep_desc")
        self._struct = None
        return self

    def __init__(self, edge_index, num_nodes, code=None, validate=True):
        nat.require_cuda(edge_index, "edge_index", torch.int64)
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise RuntimeError("mgv_b200: edge_index must be [2, E]")
        dev = edge_index.device
        self.device = dev
        self.N = int(num_nodes)
        self.E = int(edge_index.size(1))
        self.edge_index = edge_index
        self.code = None
        if code is not None:
            self.code = nat.require_cuda(code.to(torch.int32).contiguous(), "code", torch.int32)
            if self.code.numel() != self.N:
                raise RuntimeError("mgv_b200: code must have one entry per node")
        i32 = dict(dtype=torch.int32, device=dev)
        self.in_ptr = torch.empty(self.N + 1, **i32)
        self.in_src = torch.empty(max(self.E, 1), **i32)
        self.out_ptr = torch.empty(self.N + 1, **i32)
        self.out_pack = torch.empty(max(self.E, 1), **i32)
        self.out_slot = torch.empty(max(self.E, 1), **i32)
        lib = nat.lib()
        with nat.on_device(dev):
            nb = lib.mgv_csr_workspace_bytes(self.N, self.E)
            ws = nat.workspace(nb, dev)
            nat.check(lib.mgv_build_csr(nat.ptr(edge_index), self.E, self.N, nat.ptr(self.code),
                                        nat.ptr(self.in_ptr), nat.ptr(self.in_src), nat.ptr(self.out_ptr),
                                        nat.ptr(self.out_pack), nat.ptr(self.out_slot), nat.ptr(ws), nb,
                                        None if validate else nat.ptr(error_word(dev)),
                                        nat.stream_of(dev)), "mgv_build_csr")
            # degree orders + tile cost prefixes for the tensor-core tiles of the struct encoder
            ntiles = (self.N + nat.TILE_ROWS - 1) // nat.TILE_ROWS
            self.deg_order_in = torch.empty(max(self.N, 1), **i32)
            self.deg_order_out = torch.empty(max(self.N, 1), **i32)
            self.tile_cost_in = torch.empty(ntiles + 1, **i32)
            self.tile_cost_out = torch.empty(ntiles + 1, **i32)
            self.gdesc_in = torch.empty(max(self.N, 1), 4, **i32)
            self.gdesc_out = torch.empty(max(self.N, 1), 4, **i32)
            nb = lib.mgv_degree_order_workspace_bytes(self.N)
            ws = nat.workspace(nb, dev)
            for p_, i_, o_, g_, c_ in ((self.in_ptr, self.in_src, self.deg_order_in, self.gdesc_in, self.tile_cost_in),
                                       (self.out_ptr, self.out_pack, self.deg_order_out, self.gdesc_out, self.tile_cost_out)):
                nat.check(lib.mgv_build_degree_order(nat.ptr(p_), nat.ptr(i_), self.N, nat.ptr(o_), nat.ptr(g_), nat.ptr(c_),
                                                     nat.ptr(ws), nb, nat.stream_of(dev)), "mgv_build_degree_order")
        self.level = None
        self.L = 1
        self.streams = 1
        self.order = None
        self.seg_ptr = None
        self.sweep_desc = None
        self.code_count = [0] * nat.NCODE
        self._struct = None

    # ------------------------------------------------------------------ levels
    def levelize(self, reverse=False):
        """ASAP levels (== top_sort).  Returns (int32 level tensor, number of levels)."""
        lib = nat.lib()
        level = torch.empty(max(self.N, 1), dtype=torch.int32, device=self.device)
        info = (ctypes.c_int32 * 2)()
        a_ptr, b_ptr, b_idx = ((self.in_ptr, self.out_ptr, self.out_pack) if not reverse
                               else (self.out_ptr, self.in_ptr, self.in_src))
        with torch.cuda.device(self.device):
            nb = lib.mgv_levelize_workspace_bytes(self.N)
            ws = nat.workspace(nb, self.device)
            nat.check(lib.mgv_levelize(nat.ptr(a_ptr), nat.ptr(b_ptr), nat.ptr(b_idx), self.N, nat.ptr(level),
                                       info, nat.ptr(ws), nb, nat.stream_of(self.device)), "mgv_levelize")
        return level[:self.N], max(int(info[0]), 1)

    def set_levels(self, level=None, num_levels=None, code_count=None, stream_of_node=None, streams=1):
        """Attach levels (given, e.g. ``G.forward_level``, or computed) and build the
        (stream, level, code)-segmented node lists.  With ``num_levels`` and ``code_count`` supplied (host metadata of the
        batch, data.attach_schedule_meta) nothing here synchronises with the device.  ``stream_of_node`` (int32 [N] in
        [0, streams)) cuts the batch into independent node sets (whole circuits), see data.attach_streams."""
        if self.code is None:
            raise RuntimeError("mgv_b200: a level schedule needs gate codes")
        if level is None:
            lvl, L = self.levelize()
        else:
            lvl = nat.require_cuda(level.reshape(-1).to(torch.int32).contiguous(), "forward_level", torch.int32)
            if lvl.numel() != self.N:
                raise RuntimeError("mgv_b200: forward_level must have one entry per node")
            if num_levels is not None:
                L = max(int(num_levels), 1)
            else:
                # one device->host sync, as the reference's max(G.forward_level).item() (dg_ae_model_mig.py:67)
                L = int(lvl.max().item()) + 1 if self.N > 0 else 1
        self.level, self.L = lvl, L
        streams = int(streams) if stream_of_node is not None else 1
        if streams > 1:
            stream_of_node = nat.require_cuda(stream_of_node.reshape(-1).to(torch.int32).contiguous(), "sweep_stream", torch.int32)
            if stream_of_node.numel() != self.N:
                raise RuntimeError("mgv_b200: sweep_stream must have one entry per node")
        else:
            stream_of_node, streams = None, 1
        self.streams = streams
        i32 = dict(dtype=torch.int32, device=self.device)
        self.order = torch.empty(max(self.N, 1), **i32)
        self.seg_ptr = torch.empty(streams * L * nat.NCODE + 1, **i32)
        use_async = code_count is not None and level is not None and num_levels is not None
        counts = (ctypes.c_int64 * nat.NCODE)()
        lib = nat.lib()
        with torch.cuda.device(self.device):
            nb = lib.mgv_level_lists_workspace_bytes(self.N, streams * L)
            ws = nat.workspace(nb, self.device)
            nat.check(lib.mgv_build_level_lists(nat.ptr(lvl), nat.ptr(self.code), nat.ptr(stream_of_node), streams,
                                                self.N, L, nat.ptr(self.order),
                                                nat.ptr(self.seg_ptr), None if use_async else counts, nat.ptr(ws), nb,
                                                nat.ptr(error_word(self.device)) if use_async else None,
                                                nat.stream_of(self.device)), "mgv_build_level_lists")
            # row descriptors of the level sweep (one coalesced load per tile row instead of the index chain)
            self.sweep_desc = torch.empty(max(self.N, 1), 8, **i32)
            nat.check(lib.mgv_build_sweep_desc(nat.ptr(self.order), nat.ptr(self.in_ptr), nat.ptr(self.in_src), nat.ptr(self.out_ptr),
                                               self.N, nat.ptr(self.sweep_desc), nat.stream_of(self.device)), "mgv_build_sweep_desc")
        self.code_count = [int(c) for c in (code_count if use_async else counts)]
        self._struct = None
        return self

    # ------------------------------------------------------------------ C view
    def c_struct(self):
        if self._struct is None:
            s = nat.mgv_schedule()
            s.N, s.L, s.E = self.N, self.L, self.E
            s.order, s.seg_ptr = nat.ptr(self.order), nat.ptr(self.seg_ptr)
            s.in_ptr, s.in_src = nat.ptr(self.in_ptr), nat.ptr(self.in_src)
            s.out_ptr, s.out_pack, s.out_slot = nat.ptr(self.out_ptr), nat.ptr(self.out_pack), nat.ptr(self.out_slot)
            for c in range(nat.NCODE):
                s.code_count[c] = self.code_count[c]
            s.deg_order_in, s.deg_order_out = nat.ptr(self.deg_order_in), nat.ptr(self.deg_order_out)
            s.tile_cost_in, s.tile_cost_out = nat.ptr(self.tile_cost_in), nat.ptr(self.tile_cost_out)
            s.gdesc_in, s.gdesc_out = nat.ptr(self.gdesc_in), nat.ptr(self.gdesc_out)
            s.streams, s.reserved0 = self.streams, 0
            s.sweep_desc = nat.ptr(self.sweep_desc)
            self._struct = s
        return ctypes.byref(self._struct)

    def num_propagated(self, handled_codes):
        """Nodes the sweep updates per round: level >= 1 and a handled gate code."""
        return sum(self.code_count[c] for c in handled_codes)


_last = {"key": None, "csr": None}


def _key(edge_index, num_nodes):
    return (edge_index.data_ptr(), tuple(edge_index.shape), int(num_nodes), edge_index._version, str(edge_index.device))


def csr_for(edge_index, num_nodes):
    """CSR of ``edge_index`` with a one-entry cache, so the struct encoder called through its
    public ``forward(s, t, edge_index)`` reuses the CSR the model just built for the batch."""
    k = _key(edge_index, num_nodes)
    if _last["key"] == k:
        return _last["csr"]
    csr = GraphCSR(edge_index.contiguous(), num_nodes)
    _last["key"], _last["csr"] = k, csr
    return csr


def schedule_for_batch(G, streams=None):
    """Level schedule of a batch ``G`` (cached on the object).  Uses ``G.forward_level`` when the
    batch carries it (the reference computes it at dataset build, parser_func_others.py:63) and the batch's circuit
    sets ``G.sweep_stream`` (data.attach_streams) unless ``streams=1`` asks for plain (level, code) lists."""
    sch = getattr(G, "_mgv_schedule", None)
    ei = G.edge_index
    n = int(G.gate.shape[0])
    sos = getattr(G, "sweep_stream", None)
    want = int(getattr(G, "sweep_streams", 1)) if (sos is not None and sos.numel() == n and sos.is_cuda) else 1
    if want > 1 and streams is not None and int(streams) < want:
        want = 1                                               # the cut cannot be coarsened here: plain (level, code) lists
    if (sch is not None and sch.N == n and sch.edge_index.data_ptr() == ei.data_ptr() and sch.E == ei.size(1)
            and sch.streams == want):
        return sch
    if not ei.is_cuda:
        raise RuntimeError("mgv_b200: the batch must be on a CUDA device (no CPU path); call batch.to('cuda')")
    host_sched = getattr(G, "sched_order", None)
    if (host_sched is not None and host_sched.is_cuda and int(getattr(G, "sched_streams", 1)) == want
            and getattr(G, "num_levels", None) is not None and not os.environ.get("MGV_DEVICE_SCHEDULE")):
        sch = GraphCSR.from_host_schedule(G)
        try:
            G._mgv_schedule = sch
        except Exception:
            pass
        _last["key"], _last["csr"] = _key(ei, n), sch
        return sch
    code = G.gate.reshape(-1)
    level = getattr(G, "forward_level", None)
    level = level if (level is not None and level.numel() == n) else None
    L, counts = getattr(G, "num_levels", None), getattr(G, "level_code_count", None)
    has_meta = level is not None and L is not None and counts is not None and len(counts) == nat.NCODE
    sch = GraphCSR(ei.contiguous(), n, code=code, validate=not has_meta)
    sch.set_levels(level, L if has_meta else None, counts if has_meta else None,
                   stream_of_node=sos if want > 1 else None, streams=want)
    try:
        G._mgv_schedule = sch
    except Exception:
        pass
    _last["key"], _last["csr"] = _key(ei, n), sch
    return sch
