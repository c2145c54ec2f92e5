"""Readout MLP (reference arch/mlp.py:14-56): Linear / BatchNorm1d / ReLU / Dropout stack; the layer indices inside ``fc`` fix
the checkpoint keys (readout_prob.fc.{0,1,4,5,8}.*).  The modules are torch.nn (callable stand-alone on any device, like the
reference's); in the model's readout configuration (64 -> 32 -> 32 -> 1, batchnorm, relu, dropout) ``Model.pred_prob`` and the
Trainer's probability loss run through the fused head of csrc/readout.cu instead (``fused_head_ok`` / ops.readout_head)."""
import os

import torch.nn as nn

from ..ops import Linear

_NORMS = {"batchnorm": nn.BatchNorm1d}
_ACTS = {"relu": nn.ReLU, "relu6": nn.ReLU6, "sigmoid": nn.Sigmoid}


class MLP(nn.Module):
    def __init__(self, dim_in=256, dim_hidden=32, dim_pred=1, num_layer=3, norm_layer=None, act_layer=None,
                 p_drop=0.5, sigmoid=False, tanh=False):
        super().__init__()
        assert num_layer >= 2, "The number of layers shoud be larger or equal to 2."
        widths = [dim_in] + [dim_hidden] * (num_layer - 1)
        layers = []
        for i in range(num_layer - 1):
            layers.append(Linear(widths[i], widths[i + 1]))
            if norm_layer in _NORMS:
                layers.append(_NORMS[norm_layer](widths[i + 1]))
            if act_layer in _ACTS:
                layers.append(_ACTS[act_layer](inplace=True))
            if p_drop > 0:
                layers.append(nn.Dropout(p_drop))
        layers.append(Linear(dim_hidden, dim_pred))
        if sigmoid:
            layers.append(nn.Sigmoid())
        if tanh:
            layers.append(nn.Tanh())
        self.fc = nn.Sequential(*layers)

    def fused_head_ok(self, x):
        """True when csrc/readout.cu implements exactly this stack for ``x`` (CUDA fp32 [N, 64])."""
        fc = self.fc
        return (x.is_cuda and x.dim() == 2 and x.size(1) == 64 and len(fc) == 9
                and isinstance(fc[0], nn.Linear) and fc[0].out_features == 32 and isinstance(fc[1], nn.BatchNorm1d)
                and isinstance(fc[2], nn.ReLU) and isinstance(fc[3], nn.Dropout) and isinstance(fc[4], nn.Linear)
                and fc[4].out_features == 32 and isinstance(fc[5], nn.BatchNorm1d) and isinstance(fc[6], nn.ReLU)
                and isinstance(fc[7], nn.Dropout) and isinstance(fc[8], nn.Linear) and fc[8].out_features == 1
                and fc[1].track_running_stats and fc[1].affine and fc[5].track_running_stats and fc[5].affine
                and fc[3].p == fc[7].p and not os.environ.get("MGV_READOUT_TORCH"))

    def forward(self, x):
        return self.fc(x)
