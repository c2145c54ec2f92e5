from .mlp import MLP
from .gcn_conv import AggConv
from .tfmlp import TFMlpAggr
