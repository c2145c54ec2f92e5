"""Minimal pure-torch stand-in for the parts of PyTorch-Geometric that the
reference (959AI994/Multi-Gate-VAE, DG_VAE/deepgate) imports.

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference source
under /root/reference can be imported in the build container (which has no
torch_geometric wheel) to produce the golden vectors in tests/golden/.  It is
never imported by the product package.

Semantics restated from the published PyG 2.x behaviour that the reference's
call sites rely on (arch/tfmlp.py:35-46, arch/gcn_conv.py:34-42,
dg_ae_model_mig.py:177-180, parser_func_others.py:10-40).
"""
__version__ = "2.3.0-shim"


def is_debug_enabled():
    return False
