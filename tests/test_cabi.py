"""CPU-side checks of the C-ABI boundary: the shared library loads and exports every symbol
include/mgv_b200.h declares (no compute calls -- no GPU needed), and the host mirror keeps the
reference's module / checkpoint surface."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "mgv_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgv_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from deepgate import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_native.LIB_PATH)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(handle, n), "libmgv_b200.so does not export %s" % n
    assert set(names) == set(_native.EXPORTED_SYMBOLS)
    assert _native.lib().mgv_version() >= 100


def test_header_constants_match_binding():
    from deepgate import _native
    txt = open(os.path.join(ROOT, "include", "mgv_b200.h")).read()
    consts = dict(re.findall(r"#define (MGV_[A-Z_]+) (\d+)", txt))
    assert int(consts["MGV_D"]) == _native.D and int(consts["MGV_NCODE"]) == _native.NCODE
    assert int(consts["MGV_SWEEP_PACK_FLOATS"]) == _native.SWEEP_PACK_FLOATS
    assert int(consts["MGV_SWEEP_GRAD_FLOATS"]) == _native.SWEEP_GRAD_FLOATS
    assert int(consts["MGV_STRUCT_PACK_FLOATS"]) == _native.STRUCT_PACK_FLOATS
    assert int(consts["MGV_STRUCT_GRAD_FLOATS"]) == _native.STRUCT_GRAD_FLOATS


@pytest.mark.parametrize("kind,count", [("aig", 240931), ("mig", 340645), ("xmg", 390502), ("xag", 290788)])
def test_state_dict_surface_matches_reference(kind, count):
    """Same parameter names / shapes as the reference models."""
    import deepgate
    from oracle import dg_oracle as O
    from util import KIND_MODULE
    enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=4, t_rounds=4,
                                                     layernorm=True)
    m = getattr(deepgate, KIND_MODULE[kind]).Model(struct_encoder=enc, dim_hidden=64)
    assert sum(p.numel() for p in m.parameters()) == count
    own = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    assert own == {k: tuple(s) for k, s in O.param_shapes(kind).items()}
    assert deepgate.Model is deepgate.dg_ae_model_xag.Model          # deepgate/__init__.py:1-4 quirk


def test_product_path_refuses_cpu():
    import deepgate
    from deepgate import synth
    enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=1, t_rounds=1)
    m = deepgate.dg_ae_model_mig.Model(struct_encoder=enc, dim_hidden=64)
    b = deepgate.circuits_to_batch(synth.make_circuits("mig", 1, 4, 10, cfg=9))
    with pytest.raises(RuntimeError, match="no CPU"):
        m(b)


def test_host_levelisation_and_collate_match_oracle():
    import deepgate
    from deepgate import synth
    from deepgate.utils.dag_utils import top_sort_host
    from oracle import dg_oracle as O
    circuits = synth.make_circuits("xmg", 3, 6, 80, cfg=5, window=10)
    b = deepgate.circuits_to_batch(circuits)
    n = b.x.size(0)
    assert torch.equal(b.forward_level, O.top_sort(b.edge_index, n))           # levels shared across circuits
    assert torch.equal(b.backward_level, O.top_sort(b.edge_index.flip(0), n))
    assert torch.equal(top_sort_host(b.edge_index.numpy(), n), b.forward_level)
    off = 0
    for c in circuits:                                                           # 'index' keys are shifted
        e = c["edge_index"].shape[0]
        off += c["x"].shape[0]
    assert int(b.edge_index.max()) < n and int(b.tt_pair_index.max()) < n and b.edge_index.size(1) == sum(
        c["edge_index"].shape[0] for c in circuits)
    with pytest.raises(ValueError):
        top_sort_host(torch.tensor([[0, 1], [1, 0]]).numpy(), 2)


def test_rank_batches_are_size_balanced_and_rank_zero_is_the_single_gpu_workload():
    """bench.py draws, for every rank, different circuits of the SAME sizes (size_cfg); rank 0 of an N-GPU run is the
    single-GPU workload."""
    import numpy as np
    from deepgate import synth
    base = synth.make_circuits("aig", 6, (16, 64), (500, 1500), cfg=2)
    r0 = synth.make_circuits("aig", 6, (16, 64), (500, 1500), cfg=2, size_cfg=2)
    r3 = synth.make_circuits("aig", 6, (16, 64), (500, 1500), cfg=302, size_cfg=2)
    for a, b, c in zip(base, r0, r3):
        assert np.array_equal(a["edge_index"], b["edge_index"]) and np.array_equal(a["x"], b["x"])
        assert a["x"].shape == c["x"].shape                       # same node count ...
        assert not np.array_equal(a["edge_index"], c["edge_index"])   # ... different circuit
