"""Readout MLP (reference arch/mlp.py:14-56): Linear / BatchNorm1d / ReLU / Dropout stack.
torch.nn modules (adjacent to the hot path, SURVEY.md section 8 a11 / f#3) except that the Linear layers take their
weight gradient from csrc/linear.cu (ops.Linear: the node-dimension reduction is one SM's work for a library GEMM); the
layer indices inside ``fc`` fix the checkpoint keys (readout_prob.fc.{0,1,4,5,8}.*)."""
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

    def forward(self, x):
        return self.fc(x)
