"""Per-circuit parsing into ``OrderedData`` (reference parser_func_others.py:43-78 and, for the
AIG on-disk layout, parser_func.py:43-69).  The level schedule attached here is what
``Model.forward`` later reads as ``G.forward_level``."""
import numpy as np
import torch

from .data import OrderedData
from .utils.dag_utils import top_sort_host


def _one_hot_codes(x, num_gate_types):
    code = torch.as_tensor(np.asarray(x)[:, 1].astype(np.int64))
    return torch.nn.functional.one_hot(code, num_gate_types).to(torch.float32)


def parse_pyg_mlpgate(x, edge_index, y, tt_sim, tt_pair_index, num_gate_types=6, transposed=True):
    """``transposed=True``: edge_index [E,2] / tt_pair_index [P,2] on disk (MIG/XMG/XAG);
    ``False``: already [2,E] / [2,P] (AIG, parser_func.py)."""
    x = np.asarray(x)
    n = len(x)
    x_torch = _one_hot_codes(x, num_gate_types)
    pairs = torch.as_tensor(np.asarray(tt_pair_index), dtype=torch.long)
    ei = torch.as_tensor(np.asarray(edge_index), dtype=torch.long)
    if transposed:
        pairs, ei = pairs.t().contiguous(), ei.t().contiguous()
    idx = torch.arange(n, dtype=torch.long)
    if ei.numel() == 0:
        ei = ei.reshape(2, 0)
        fl = bl = torch.zeros(n, dtype=torch.long)
    else:
        fl = top_sort_host(ei.numpy(), n)
        bl = top_sort_host(ei.flip(0).numpy(), n)
    g = OrderedData(x=x_torch, edge_index=ei, tt_pair_index=pairs, tt_sim=torch.as_tensor(np.asarray(tt_sim)),
                    forward_level=fl, forward_index=idx, backward_level=bl, backward_index=idx.clone())
    g.gate = torch.as_tensor(x[:, 1:2].astype(np.float32))
    g.prob = torch.as_tensor(np.asarray(y, dtype=np.float32)).reshape(n, 1)
    attach_circuit_csr(g)
    return g


def attach_circuit_csr(g):
    """In / out-edge CSR of ONE circuit in local node and edge ids, computed here -- where the reference computes
    ``forward_level`` (parser_func_others.py:63) -- so that a batch's CSR is a concatenation with offsets at collate time
    (data.attach_host_schedule) instead of sorts on the device in every step.  Within a node, edges keep their original
    order (the order ``subgraph`` concatenates them in, utils/dag_utils.py:91-105)."""
    n = int(g.x.size(0))
    ei = g.edge_index.numpy()
    src, dst = ei[0], ei[1]
    if src.size and (min(src.min(), dst.min()) < 0 or max(src.max(), dst.max()) >= n):
        raise IndexError("edge_index refers to a node outside [0, %d)" % n)
    by_dst = np.argsort(dst, kind="stable")
    by_src = np.argsort(src, kind="stable")
    pos_in = np.empty(src.size, dtype=np.int64)
    pos_in[by_dst] = np.arange(src.size)
    g.csr_in_src = torch.from_numpy(src[by_dst].astype(np.int32))
    g.csr_in_deg = torch.from_numpy(np.bincount(dst, minlength=n).astype(np.int32))
    g.csr_out_dst = torch.from_numpy(dst[by_src].astype(np.int32))
    g.csr_out_slot = torch.from_numpy(pos_in[by_src].astype(np.int32))
    g.csr_out_deg = torch.from_numpy(np.bincount(src, minlength=n).astype(np.int32))
    return g


def circuits_to_batch(circuits, device=None):
    """Synthetic / parsed circuit dicts (see synth.py) -> one collated batch."""
    from .data import collate
    graphs = [parse_pyg_mlpgate(c["x"], c["edge_index"], c["prob"], c["tt_sim"], c["tt_pair_index"]) for c in circuits]
    batch = collate(graphs)
    return batch.to(device) if device is not None else batch
