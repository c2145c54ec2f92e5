"""torch_geometric.utils subset: softmax, degree, self-loop helpers,
negative_sampling, to_undirected (pure torch)."""
import torch


def _group_max(src, index, num_nodes):
    out = src.new_full((num_nodes,) + tuple(src.shape[1:]), float("-inf"))
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    # PyG: out = exp(src - max_group(src).detach()); out / (sum_group(out) + 1e-16)
    assert index is not None and dim == 0
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    src_max = _group_max(src.detach(), index, n)
    out = (src - src_max.index_select(0, index)).exp()
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(out)
    out_sum = torch.zeros((n,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    out_sum = out_sum.scatter_add(0, idx, out) + 1e-16
    return out / out_sum.index_select(0, index)


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), edge_attr


def to_undirected(edge_index, num_nodes=None):
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    return torch.unique(both, dim=1)


def negative_sampling(edge_index, num_nodes=None, num_neg_samples=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    k = edge_index.size(1) if num_neg_samples is None else int(num_neg_samples)
    taken = set((edge_index[0] * n + edge_index[1]).tolist())
    out = []
    g = torch.Generator().manual_seed(0)
    while len(out) < k and len(out) + len(taken) < n * n:
        cand = torch.randint(0, n * n, (2 * (k - len(out)) + 8,), generator=g).tolist()
        for c in cand:
            if c not in taken:
                taken.add(c)
                out.append(c)
                if len(out) == k:
                    break
    out = torch.tensor(out, dtype=torch.long, device=edge_index.device)
    return torch.stack([out // n, out % n], dim=0)
