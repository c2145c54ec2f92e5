"""Level schedule helpers with the reference's names (utils/dag_utils.py:10-37, 80-105).

``top_sort`` / ``return_order_info`` / ``subgraph`` are answered from the device-built CSR
(csrc/schedule.cu) and are bit-exact with the reference's host loops.  ``top_sort_host`` is the
O(N+E) numpy Kahn peel used when a dataset is parsed on a machine without a GPU.
"""
import numpy as np
import torch

from ..schedule import GraphCSR


def _as_cuda_edge_index(edge_index):
    ei = torch.as_tensor(edge_index)
    if not ei.is_cuda:
        ei = ei.to("cuda")
    return ei.to(torch.int64).contiguous()


def top_sort(edge_index, graph_size):
    """ASAP level of every node, int64 [graph_size] on the device of ``edge_index`` (CUDA)."""
    ei = _as_cuda_edge_index(edge_index)
    level, _ = GraphCSR(ei, graph_size).levelize()
    return level.to(torch.int64)


def return_order_info(edge_index, num_nodes):
    ei = _as_cuda_edge_index(edge_index)
    csr = GraphCSR(ei, num_nodes)
    fl, _ = csr.levelize()
    bl, _ = csr.levelize(reverse=True)
    idx = torch.arange(num_nodes, dtype=torch.int64, device=ei.device)
    return fl.to(torch.int64), idx, bl.to(torch.int64), idx.clone()


def subgraph(target_idx, edge_index, edge_attr=None, dim=1):
    """Incoming (dim=1) / outgoing (dim=0) edges of ``target_idx``, grouped per target in the given
    order, ascending edge id inside a target (reference dag_utils.py:91-105)."""
    ei = _as_cuda_edge_index(edge_index)
    n = int(max(int(ei.max()) if ei.numel() else -1, int(target_idx.max()) if len(target_idx) else -1)) + 1
    csr = GraphCSR(ei, n)
    tgt = torch.as_tensor(target_idx, device=ei.device).to(torch.int64)
    ptr, other = (csr.in_ptr, csr.in_src) if dim == 1 else (csr.out_ptr, csr.out_pack)
    beg, end = ptr[tgt].long(), ptr[tgt + 1].long()
    cnt = end - beg
    owner = torch.repeat_interleave(torch.arange(tgt.numel(), device=ei.device), cnt)
    offs = torch.arange(int(cnt.sum()), device=ei.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
    slot = beg[owner] + offs
    oth = (other[slot].long() & ((1 << 28) - 1))
    me = tgt[owner]
    sub = torch.stack([oth, me]) if dim == 1 else torch.stack([me, oth])
    if edge_attr is not None:
        raise NotImplementedError("mgv_b200: edge attributes are not used by any live model")
    return sub, None


def top_sort_host(edge_index, graph_size):
    """numpy Kahn levelisation (data-pipeline side, no GPU needed); same result as top_sort."""
    ei = np.asarray(edge_index)
    src, dst = ei[0].astype(np.int64), ei[1].astype(np.int64)
    indeg = np.bincount(dst, minlength=graph_size).astype(np.int64)
    order = np.argsort(src, kind="stable")
    out_ptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=graph_size))])
    out_dst = dst[order]
    level = np.zeros(graph_size, dtype=np.int64)
    frontier = np.nonzero(indeg == 0)[0]
    done, lv = 0, 0
    while frontier.size:
        done += frontier.size
        cnt = out_ptr[frontier + 1] - out_ptr[frontier]
        starts = np.repeat(out_ptr[frontier], cnt)
        offs = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        hit = out_dst[starts + offs]
        np.subtract.at(indeg, hit, 1)
        nxt = np.unique(hit[indeg[hit] == 0])
        lv += 1
        level[nxt] = lv
        frontier = nxt
    if done != graph_size:
        raise ValueError("cycle in circuit graph")
    return torch.from_numpy(level)
