"""DG_AE model, AIG: AND = 1, NOT = 2 (reference dg_ae_model_aig.py:26-100)."""
from .dg_ae_model_base import LevelModel


class Model(LevelModel):
    ENCODER_ATTR = "struct_encoder"
    GATE_MODULES = ((1, "and"), (2, "not"))
