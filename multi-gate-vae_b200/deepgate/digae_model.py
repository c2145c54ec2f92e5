"""Graph auto-encoder wrappers (reference digae_model.py:26-169): ``GAE`` (undirected inner product) and
``DirectedGAE`` (source / target embeddings, directed inner-product decoder).

``train.py:10`` imports this module; the live models (dg_ae_model_*.py) carry their own copy of ``recon_loss``.  Same
constructor arguments and method names as upstream.  ``DirectedGAE.recon_loss`` runs the fused decoder + BCE kernel and
the device-side negative sampler (csrc/recon.cu) like ``LevelModel.recon_loss``; there is no PyG dependency.
"""
import torch

from . import ops
from .digae_layer import DirectedInnerProductDecoder

EPS = 1e-15
MAX_LOGSTD = 10


def reset(nn):
    """Re-initialise a module or each of its children (digae_model.py:14-24)."""
    if nn is None:
        return
    children = list(nn.children()) if hasattr(nn, "children") else []
    for item in (children if children else [nn]):
        if hasattr(item, "reset_parameters"):
            item.reset_parameters()


class InnerProductDecoder(torch.nn.Module):
    """sigma(<z[src], z[dst]>), the decoder ``GAE`` defaults to."""

    def forward(self, z, edge_index, sigmoid=True):
        value = (z[edge_index[0]] * z[edge_index[1]]).sum(dim=1)
        return torch.sigmoid(value) if sigmoid else value

    def forward_all(self, z, sigmoid=True):
        adj = z @ z.t()
        return torch.sigmoid(adj) if sigmoid else adj


def _auc_ap(pos_pred, neg_pred):
    from sklearn.metrics import average_precision_score, roc_auc_score
    y = torch.cat([torch.ones_like(pos_pred), torch.zeros_like(neg_pred)]).detach().cpu().numpy()
    pred = torch.cat([pos_pred, neg_pred]).detach().cpu().numpy()
    return roc_auc_score(y, pred), average_precision_score(y, pred)


def _negatives(pos_edge_index, num_nodes):
    """Stand-in for ``negative_sampling(add_self_loops(remove_self_loops(pos)), N)``: one negative per entry of that edge
    set, none of them a self loop or an edge."""
    from .schedule import csr_for
    keep = pos_edge_index[0] != pos_edge_index[1]
    csr = csr_for(pos_edge_index, num_nodes)
    return ops.negative_sample(csr, int(keep.sum()) + num_nodes)


class GAE(torch.nn.Module):
    def __init__(self, encoder, decoder=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = InnerProductDecoder() if decoder is None else decoder
        GAE.reset_parameters(self)

    def reset_parameters(self):
        reset(self.encoder)
        reset(self.decoder)

    def encode(self, *args, **kwargs):
        return self.encoder(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.decoder(*args, **kwargs)

    def recon_loss(self, z, pos_edge_index, neg_edge_index=None):
        pos_loss = -torch.log(self.decoder(z, pos_edge_index, sigmoid=True) + EPS).mean()
        if neg_edge_index is None:
            neg_edge_index = _negatives(pos_edge_index, z.size(0))
        neg_loss = -torch.log(1 - self.decoder(z, neg_edge_index, sigmoid=True) + EPS).mean()
        return pos_loss + neg_loss

    def test(self, z, pos_edge_index, neg_edge_index):
        return _auc_ap(self.decoder(z, pos_edge_index, sigmoid=True), self.decoder(z, neg_edge_index, sigmoid=True))


class DirectedGAE(torch.nn.Module):
    def __init__(self, encoder, decoder=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = DirectedInnerProductDecoder() if decoder is None else decoder
        DirectedGAE.reset_parameters(self)

    def reset_parameters(self):
        reset(self.encoder)
        reset(self.decoder)

    def forward(self, data):
        s, t = self.encoder(data.x, data.x, data.edge_index)
        return self.decoder.forward_all(s, t)

    def encode(self, *args, **kwargs):
        return self.encoder(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.decoder(*args, **kwargs)

    def recon_loss(self, s, t, pos_edge_index, neg_edge_index=None):
        """(loss, pred_bin, gt_bin) as digae_model.py:134-156, through the fused kernel."""
        if neg_edge_index is None:
            neg_edge_index = _negatives(pos_edge_index, s.size(0))
        loss, pred_bin = ops.recon_loss(torch.cat([s, t], dim=-1), pos_edge_index, neg_edge_index)
        gt_bin = torch.cat([torch.ones(pos_edge_index.size(1), dtype=torch.int32, device=s.device),
                            torch.zeros(neg_edge_index.size(1), dtype=torch.int32, device=s.device)])
        return loss, pred_bin, gt_bin

    def test(self, s, t, pos_edge_index, neg_edge_index):
        return _auc_ap(self.decoder(s, t, pos_edge_index, sigmoid=True), self.decoder(s, t, neg_edge_index, sigmoid=True))
