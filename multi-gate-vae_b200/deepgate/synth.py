"""Deterministic synthetic AIG / MIG / XMG / XAG netlists (workload generator).

The reference ships no data, so the benchmark and the parity
tests use synthetic circuits in the on-disk layout the reference's parsers
consume (parser_func_others.py:43-78): per circuit ``x[:, 1]`` = gate code,
``edge_index [E, 2]`` rows (src, dst), ``prob [N]``, ``tt_pair_index [P, 2]``,
``tt_sim [P]``.  Gate codes follow the reference's ``gate_to_index``
{INPUT:0, MAJ:1, NOT:2, AND:3, OR:4, XOR:5} (parser.py:133); AIG uses its own
codes AND=1, NOT=2 (dg_ae_model_aig.py:67-68).

Generator: ``n_pi`` INPUT nodes, then gates in topological id
order; every gate draws its fan-ins uniformly without replacement from the
previous ``window`` node ids (``None`` = all earlier nodes -> shallow circuit).
"""
import numpy as np

INPUT, MAJ, NOT, AND, OR, XOR = 0, 1, 2, 3, 4, 5

# kind -> list of (code, fan_in, probability)
GATE_MIX = {
    "aig": [(1, 2, 0.60), (2, 1, 0.40)],
    "mig": [(MAJ, 3, 0.70), (NOT, 1, 0.30)],
    "mig4": [(MAJ, 3, 0.60), (NOT, 1, 0.30), (AND, 2, 0.05), (OR, 2, 0.05)],
    "xag": [(AND, 2, 0.45), (XOR, 2, 0.25), (NOT, 1, 0.30)],
    "xmg": [(MAJ, 3, 0.40), (XOR, 2, 0.25), (NOT, 1, 0.25), (AND, 2, 0.05), (OR, 2, 0.05)],
}


def make_circuit(kind, n_pi, n_gates, seed, window=None, n_pairs=64):
    """One synthetic circuit as a dict of numpy arrays (see module docstring)."""
    assert n_pi >= 3, "need at least 3 primary inputs (MAJ fan-in)"
    rng = np.random.default_rng(seed)
    mix = GATE_MIX[kind]
    codes = np.array([m[0] for m in mix], dtype=np.int64)
    fanin = np.array([m[1] for m in mix], dtype=np.int64)
    p = np.array([m[2] for m in mix], dtype=np.float64)
    pick = rng.choice(len(mix), size=n_gates, p=p / p.sum())
    n = n_pi + n_gates
    code = np.zeros(n, dtype=np.int64)
    code[n_pi:] = codes[pick]
    d = fanin[pick]                                   # fan-in of each gate
    gid = np.arange(n_pi, n, dtype=np.int64)
    lo = np.zeros(n_gates, dtype=np.int64) if window is None else np.maximum(0, gid - int(window))
    span = gid - lo                                   # candidates: ids lo .. gid-1  (>= n_pi >= 3)
    u = rng.random((n_gates, 3))
    r1 = np.minimum((u[:, 0] * span).astype(np.int64), span - 1)
    r2 = np.minimum((u[:, 1] * (span - 1)).astype(np.int64), span - 2)
    r2 = r2 + (r2 >= r1)
    a, b = np.minimum(r1, r2), np.maximum(r1, r2)
    r3 = np.minimum((u[:, 2] * (span - 2)).astype(np.int64), span - 3)
    r3 = r3 + (r3 >= a)
    r3 = r3 + (r3 >= b)
    cand = np.stack([r1, r2, r3], axis=1) + lo[:, None]            # [G, 3]
    keep = np.arange(3)[None, :] < d[:, None]                      # first d draws
    src = cand[keep]
    dst = np.broadcast_to(gid[:, None], (n_gates, 3))[keep]
    edge_index = np.stack([src, dst], axis=1).astype(np.int64)     # [E, 2], gate-major order
    x = np.zeros((n, 2), dtype=np.int64)
    x[:, 0] = np.arange(n)
    x[:, 1] = code
    prob = rng.random(n).astype(np.float32)
    tt_pair_index = rng.integers(0, n, size=(n_pairs, 2), dtype=np.int64)
    tt_sim = rng.random(n_pairs).astype(np.float32)
    return {"x": x, "edge_index": edge_index, "prob": prob,
            "tt_pair_index": tt_pair_index, "tt_sim": tt_sim, "kind": kind}


def make_circuits(kind, batch, n_pi, n_gates, cfg=0, window=None, n_pairs=64, size_cfg=None):
    """``batch`` circuits; ``n_pi``/``n_gates`` may be ints or (lo, hi) ranges.
    Seed of circuit i is 1000*cfg + i.  ``size_cfg`` (default ``cfg``) seeds the SIZE draw
    separately: data-parallel ranks pass the same ``size_cfg`` and different ``cfg`` to get size-balanced batches
    (same circuit sizes, different circuits), the usual bucketed sampler of a DDP loader."""
    out = []
    size_cfg = cfg if size_cfg is None else size_cfg
    for i in range(batch):
        seed = 1000 * cfg + i
        r = np.random.default_rng(10_000_019 * (size_cfg + 1) + i)
        pi = n_pi if np.isscalar(n_pi) else int(r.integers(n_pi[0], n_pi[1] + 1))
        ng = n_gates if np.isscalar(n_gates) else int(r.integers(n_gates[0], n_gates[1] + 1))
        out.append(make_circuit(kind, pi, ng, seed, window=window, n_pairs=n_pairs))
    return out
