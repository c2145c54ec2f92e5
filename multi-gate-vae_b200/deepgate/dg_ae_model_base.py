"""Shared implementation of the four DG_AE models (reference dg_ae_model_{aig,mig,xmg,xag}.py).

A subclass only states (a) the attribute name of its struct encoder, (b) its gate-code ->
aggregator/GRU map in the reference's module registration order -- those two things fix the
checkpoint keys.  ``forward(G)`` is:
    level schedule (device-built, cached on the batch)  ->  struct encoder kernels (s, t)
    -> hs = hs_linear([s || t])  ->  level sweep kernel (all rounds, all levels)  ->  (hs, hf).
"""
import os

import torch
from torch import nn

from . import ops
from .arch.mlp import MLP
from .arch.tfmlp import TFMlpAggr
from .digae_layer import DirectedInnerProductDecoder
from .schedule import schedule_for_batch

EPS = 1e-15
MAX_LOGSTD = 10


class LevelModel(nn.Module):
    """Recurrent GNN over circuit levels with structural (hs) and functional (hf) states."""

    ENCODER_ATTR = "struct_encoder"
    GATE_MODULES = ()          # ((gate code, suffix), ...) in registration order

    def __init__(self, struct_encoder, num_rounds=1, dim_hidden=128, enable_encode=True, enable_reverse=True):
        super().__init__()
        setattr(self, self.ENCODER_ATTR, struct_encoder)
        self.decoder = DirectedInnerProductDecoder()
        self.hs_linear = ops.Linear(dim_hidden * 2, dim_hidden)
        self.hs_decompose = ops.Linear(dim_hidden, dim_hidden * 2)
        self.num_rounds = num_rounds
        self.enable_encode = enable_encode
        self.enable_reverse = enable_reverse
        self.dim_hidden = dim_hidden
        self.dim_mlp = 32
        for _, suffix in self.GATE_MODULES:
            setattr(self, "aggr_%s_func" % suffix, TFMlpAggr(dim_hidden * 2, dim_hidden))
        for _, suffix in self.GATE_MODULES:
            setattr(self, "update_%s_func" % suffix, nn.GRU(dim_hidden, dim_hidden))
        self.readout_prob = MLP(dim_hidden, self.dim_mlp, 1, num_layer=3, p_drop=0.2, norm_layer="batchnorm",
                                act_layer="relu")

    # ------------------------------------------------------------------ hot path
    def forward(self, G):
        sched = schedule_for_batch(G)
        encoder = getattr(self, self.ENCODER_ATTR)
        # the reference feeds one_hot(G.x[:, 1], 6) although G.x is already one-hot, i.e. the
        # feature is one_hot(1{gate code == 1}, 6)  (dg_ae_model_mig.py:71)
        feat = torch.nn.functional.one_hot(G.x[:, 1].to(torch.int64), num_classes=6).to(torch.float32)
        s, t = encoder(feat, feat, G.edge_index)
        hs = self.hs_linear(torch.cat([s, t], dim=-1))
        codes = [c for c, _ in self.GATE_MODULES]
        modules = [(getattr(self, "aggr_%s_func" % sfx), getattr(self, "update_%s_func" % sfx))
                   for _, sfx in self.GATE_MODULES]
        hf = ops.level_sweep(hs, sched, self.num_rounds, codes, modules)
        return hs, hf

    # ------------------------------------------------------------------ heads (torch.nn, adjacent to the path)
    def pred_prob(self, hf):
        return torch.clamp(self.readout_prob(hf), min=0.0, max=1.0)

    def recon_loss(self, hs, pos_edge_index, neg_edge_index=None):
        """dg_ae_model_mig.py:169-191.  The decoder + BCE terms run in one fused kernel (ops.recon_loss); missing
        negatives are drawn on the device (no self loops, no existing edges) without a host sync."""
        st = self.hs_decompose(hs)
        if neg_edge_index is None:
            from . import schedule
            csr = schedule._last["csr"]
            if csr is None or csr.N != st.size(0) or csr.E != pos_edge_index.size(1) or csr.device != st.device:
                csr = schedule.csr_for(pos_edge_index, st.size(0))
            neg_edge_index = ops.negative_sample(csr, pos_edge_index.size(1))
        loss, pred_bin = ops.recon_loss(st, pos_edge_index, neg_edge_index)
        ep, en = pos_edge_index.size(1), neg_edge_index.size(1)
        gt_bin = torch.cat([torch.ones(ep, dtype=torch.int32, device=hs.device),
                            torch.zeros(en, dtype=torch.int32, device=hs.device)])
        return loss, pred_bin, gt_bin

    # ------------------------------------------------------------------ checkpoints
    def load(self, model_path):
        checkpoint = torch.load(model_path, map_location=lambda storage, loc: storage)
        loaded = {}
        for k, v in checkpoint["state_dict"].items():
            loaded[k[7:] if (k.startswith("module") and not k.startswith("module_list")) else k] = v
        own = self.state_dict()
        for k in list(loaded):
            if k not in own:
                print("Drop parameter {}.".format(k))
            elif loaded[k].shape != own[k].shape:
                print("Skip loading parameter {}, required shape{}, loaded shape{}.".format(
                    k, own[k].shape, loaded[k].shape))
                loaded[k] = own[k]
        for k in own:
            if k not in loaded:
                print("No param {}.".format(k))
                loaded[k] = own[k]
        self.load_state_dict(loaded, strict=False)

    def load_pretrained(self, pretrained_model_path=""):
        if pretrained_model_path == "":
            pretrained_model_path = os.path.join(os.path.dirname(__file__), "pretrained", "model.pth")
        self.load(pretrained_model_path)
