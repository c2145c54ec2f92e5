"""DG_AE model, MIG: MAJ = 1, NOT = 2, AND = 3, OR = 4 (reference dg_ae_model_mig.py:21-132)."""
from .dg_ae_model_base import LevelModel


class Model(LevelModel):
    ENCODER_ATTR = "mig_struct_encoder"
    GATE_MODULES = ((3, "and"), (2, "not"), (4, "or"), (1, "maj"))
