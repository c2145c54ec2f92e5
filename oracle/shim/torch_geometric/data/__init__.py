"""torch_geometric.data subset: Data container, collate honouring
__inc__/__cat_dim__ (adds `batch`), InMemoryDataset placeholder."""
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kw):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kw.items():
            setattr(self, k, v)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return k in self.__dict__ and self.__dict__[k] is not None

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    @property
    def num_nodes(self):
        if getattr(self, "x", None) is not None:
            return self.x.size(0)
        return int(self.edge_index.max()) + 1

    def __inc__(self, key, value, *a, **k):
        return self.num_nodes if "index" in key or "face" in key else 0

    def __cat_dim__(self, key, value, *a, **k):
        return -1 if "index" in key or "face" in key else 0

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self


def collate(data_list):
    first = data_list[0]
    out = first.__class__()
    keys = [k for k in first.keys]
    offset = 0
    cols = {k: [] for k in keys}
    batch_vec = []
    for gi, d in enumerate(data_list):
        n = d.num_nodes
        for k in keys:
            v = d[k]
            if torch.is_tensor(v):
                inc = d.__inc__(k, v)
                cols[k].append(v + offset if inc else v)
            else:
                cols[k].append(v)
        batch_vec.append(torch.full((n,), gi, dtype=torch.long))
        offset += n
    for k in keys:
        if torch.is_tensor(cols[k][0]):
            out[k] = torch.cat(cols[k], dim=first.__cat_dim__(k, cols[k][0]))
        else:
            out[k] = cols[k][0]
    out.batch = torch.cat(batch_vec)
    return out


Batch = Data


class InMemoryDataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root
