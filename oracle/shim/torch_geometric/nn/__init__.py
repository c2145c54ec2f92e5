"""torch_geometric.nn subset: MessagePassing with aggr='add' only."""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert aggr == "add"
        assert flow in ("source_to_target", "target_to_source")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs.get("x")
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        n = x.size(self.node_dim)
        want = inspect.signature(self.message).parameters
        args = {}
        for name in want:
            if name == "x_i":
                args[name] = x.index_select(self.node_dim, edge_index[i])
            elif name == "x_j":
                args[name] = x.index_select(self.node_dim, edge_index[j])
            elif name == "index":
                args[name] = edge_index[i]
            elif name == "ptr":
                args[name] = None
            elif name == "size_i":
                args[name] = n
            elif name in kwargs:
                args[name] = kwargs[name]
            else:
                args[name] = None
        msg = self.message(**args)
        dim = self.node_dim if self.node_dim >= 0 else msg.dim() + self.node_dim
        shape = list(msg.shape)
        shape[dim] = n
        out = torch.zeros(shape, dtype=msg.dtype, device=msg.device)
        out = out.index_add(dim, edge_index[i], msg)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
