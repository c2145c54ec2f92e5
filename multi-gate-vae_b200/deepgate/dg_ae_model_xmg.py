"""DG_AE model, XMG: MAJ = 1, NOT = 2, AND = 3, OR = 4, XOR = 5 (reference dg_ae_model_xmg.py:22-150)."""
from .dg_ae_model_base import LevelModel


class Model(LevelModel):
    ENCODER_ATTR = "xmg_struct_encoder"
    GATE_MODULES = ((3, "and"), (2, "not"), (5, "xor"), (1, "maj"), (4, "or"))
