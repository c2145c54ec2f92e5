import torch
from ..data import collate


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=collate, **kw)
