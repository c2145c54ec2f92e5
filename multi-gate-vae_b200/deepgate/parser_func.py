"""AIG on-disk layout -> ``OrderedData`` (reference parser_func.py:43-69).

``graphs.npz`` of the AIG set stores ``edge_index`` as [2, E] and ``tt_pair_index`` as [2, P] already, so nothing is
transposed here (the MIG / XMG / XAG twin, parser_func_others.py, transposes [E, 2] / [P, 2]); ``gate`` is attached by
the caller from the npz (parser.py:118-119).  Shares the implementation with ``parser_func_others``.
"""
from .data import OrderedData                                   # noqa: F401  (re-exported, as upstream's star import does)
from . import parser_func_others as _others

__all__ = ["OrderedData", "parse_pyg_mlpgate"]


def parse_pyg_mlpgate(x, edge_index, y, tt_sim, tt_pair_index, num_gate_types=6):
    """One AIG circuit.  ``x`` [N, >=2] with the gate code in column 1; ``edge_index`` [2, E]; ``tt_pair_index`` [2, P]."""
    return _others.parse_pyg_mlpgate(x, edge_index, y, tt_sim, tt_pair_index, num_gate_types=num_gate_types, transposed=False)
