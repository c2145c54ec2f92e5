"""DG_AE model, XAG: NOT = 2, AND = 3, XOR = 5 (reference dg_ae_model_xag.py:22-124)."""
from .dg_ae_model_base import LevelModel


class Model(LevelModel):
    ENCODER_ATTR = "xag_struct_encoder"
    GATE_MODULES = ((3, "and"), (2, "not"), (5, "xor"))
