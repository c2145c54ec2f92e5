"""torch_scatter stand-in (the reference only imports the name, arch/gcn_conv.py:10)."""
import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    n = int(index.max()) + 1 if dim_size is None else dim_size
    shape = list(src.shape)
    shape[dim] = n
    res = torch.zeros(shape, dtype=src.dtype, device=src.device) if out is None else out
    return res.index_add(dim, index, src)
