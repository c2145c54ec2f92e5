"""TFMlpAggr parameter holder (reference arch/tfmlp.py:11-52): additive-attention aggregator
    a_j = attn_lin([msg_q(x_i) || msg_k(x_j)]),  alpha = softmax over the in-edges of i,
    out_i = sum_j alpha_j msg_v(x_j).
One instance per gate code (aggr_{and,not,or,maj,xor}_func).  The arithmetic runs inside the fused
level-sweep kernels (csrc/sweep.cu), which use the cancellation of the query term inside the
softmax; all twelve tensors still receive gradients (zeros where the
reference's gradient is mathematically zero)."""
import torch.nn as nn


class TFMlpAggr(nn.Module):
    def __init__(self, in_channels, ouput_channels=64, reverse=False, mlp_post=None):
        super().__init__()
        if ouput_channels is None:
            ouput_channels = in_channels
        assert in_channels > 0 and ouput_channels > 0, "The dimension for the DeepSetConv should be larger than 0."
        if mlp_post is not None or reverse:
            raise NotImplementedError("mgv_b200: TFMlpAggr mlp_post / reverse are not used by any live model")
        self.msg_post = None
        self.attn_lin = nn.Linear(ouput_channels + ouput_channels, 1)
        self.msg_q = nn.Linear(in_channels, ouput_channels)
        self.msg_k = nn.Linear(in_channels, ouput_channels)
        self.msg_v = nn.Linear(in_channels, ouput_channels)

    def forward(self, x, edge_index, edge_attr=None, **kwargs):
        raise NotImplementedError(
            "mgv_b200: TFMlpAggr is fused with the GRU update into the level-sweep CUDA kernel; "
            "call Model.forward(G)")
