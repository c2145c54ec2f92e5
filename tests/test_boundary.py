"""The data / entry-point boundary of the path: what the reference's ``train.py`` imports and calls.

CPU: import surface (train.py:5-17), the two on-disk npz layouts through ``NpzParser`` (parser.py:22-125; SURVEY.md
Appendix B #13, #14), and the collate rules against the REFERENCE-collated tensors stored in tests/golden/*.pt
(parser_func_others.py:28-40).  GPU: a replay of train.py:23-104's call sequence on a synthetic dataset."""
import os
import socket
import types

import numpy as np
import pytest
import torch

from util import load_golden

GOLDEN_CASES = {
    # name: (gate mix, batch, n_pi, n_gates, window, cfg)  -- oracle/make_golden.py CASES, n_pairs = 48
    "mig_b4_r1": ("mig4", 4, 16, 200, None, 1),
    "aig_b4_r1": ("aig", 4, 12, 160, None, 2),
    "xmg_b3_r2": ("xmg", 3, 10, 150, 24, 3),
    "xag_b3_r1": ("xag", 3, 8, 120, 12, 4),
}


def test_import_surface_of_train_py():
    """train.py:5-17 and the names it uses (train.py:33-73)."""
    import deepgate
    import deepgate.digae_layer, deepgate.digae_model, deepgate.digvae_model                      # noqa: E401,F401
    import deepgate.dg_ae_model_aig, deepgate.dg_ae_model_mig, deepgate.dg_ae_model_xag, deepgate.dg_ae_model_xmg   # noqa
    for name in ("NpzParser", "Trainer", "Model", "parse_pyg_mlpgate", "OrderedData"):
        assert hasattr(deepgate, name), name
    assert hasattr(deepgate.digae_model, "DirectedGAE") and hasattr(deepgate.digae_model, "GAE")
    assert hasattr(deepgate.digae_layer, "DirectMultiGCNEncoder") and hasattr(deepgate.digae_layer, "DirectedInnerProductDecoder")
    assert hasattr(deepgate.digvae_model, "DirectedGVAE")
    import deepgate.parser, deepgate.parser_func, deepgate.parser_func_others                     # noqa: E401,F401


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_collate_equals_reference_collated_golden(name):
    """``circuits_to_batch`` (parse + collate) reproduces, bit for bit, what the reference's parse_pyg_mlpgate + PyG collate
    produced for the same circuits (the golden files were written by the unmodified reference, oracle/make_golden.py)."""
    import deepgate
    from deepgate import synth
    mix, batch, n_pi, n_gates, window, cfg = GOLDEN_CASES[name]
    g = load_golden(name)
    circuits = synth.make_circuits(mix, batch, n_pi, n_gates, cfg=cfg, window=window, n_pairs=48)
    if name.startswith("aig"):         # AIG layout: [2, E] / [2, P] on disk, parsed by parser_func without transposes
        graphs = []
        for c in circuits:
            gr = deepgate.parser_func.parse_pyg_mlpgate(c["x"], c["edge_index"].T.copy(), c["prob"], c["tt_sim"], c["tt_pair_index"].T.copy())
            gr.gate = torch.tensor(c["x"][:, 1:2])
            graphs.append(gr)
        b = deepgate.collate(graphs)
    else:
        b = deepgate.circuits_to_batch(circuits)
    assert torch.equal(b.edge_index, g["edge_index"])
    assert torch.equal(b.forward_level.long(), g["forward_level"]) and torch.equal(b.backward_level.long(), g["backward_level"])
    assert torch.equal(b.tt_pair_index, g["tt_pair_index"]) and torch.equal(b.tt_sim, g["tt_sim"])
    assert torch.equal(b.prob, g["prob"]) and torch.equal(b.gate.reshape(-1).to(torch.int32), g["code"])
    n = g["code"].numel()
    assert torch.equal(b.forward_index, torch.arange(n)) and torch.equal(b.backward_index, torch.arange(n))
    assert torch.equal(b.x, torch.nn.functional.one_hot(g["code"].long(), 6).float())
    assert b.num_levels == int(g["forward_level"].max()) + 1 and b.num_graphs == batch


def write_npz_dataset(root, kind, circuits, extra_names=()):
    """Synthetic circuits in the reference's on-disk layouts."""
    os.makedirs(root, exist_ok=True)
    graphs, labels = {}, {}
    names = ["c%03d" % i for i in range(len(circuits))]
    for nme, c in zip(names, circuits):
        if kind == "aig":
            graphs[nme] = {"x": c["x"], "edge_index": c["edge_index"].T.copy(), "tt_sim": c["tt_sim"],
                           "tt_pair_index": c["tt_pair_index"].T.copy(), "prob": c["prob"], "gate": c["x"][:, 1:2].copy()}
        else:
            graphs[nme] = {"x": c["x"], "edge_index": c["edge_index"]}
            labels[nme] = {"tt_dis": c["tt_sim"], "tt_pair_index": c["tt_pair_index"], "prob": c["prob"]}
    for nme in extra_names:            # dropped by name (parser.py:90) or for having no pairs (parser.py:109-111)
        c = circuits[0]
        empty = nme.startswith("nopairs")
        pairs = np.zeros((0, 2), dtype=np.int64) if empty else c["tt_pair_index"]
        sims = np.zeros((0,), dtype=np.float32) if empty else c["tt_sim"]
        if kind == "aig":
            graphs[nme] = {"x": c["x"], "edge_index": c["edge_index"].T.copy(), "tt_sim": sims, "tt_pair_index": pairs.T.copy(),
                           "prob": c["prob"], "gate": c["x"][:, 1:2].copy()}
        else:
            graphs[nme] = {"x": c["x"], "edge_index": c["edge_index"]}
            labels[nme] = {"tt_dis": sims, "tt_pair_index": pairs, "prob": c["prob"]}
    cpath = os.path.join(root, "graphs.npz")
    np.savez(cpath, circuits=np.array(graphs, dtype=object))
    lpath = cpath
    if kind != "aig":
        lpath = os.path.join(root, "labels.npz")
        np.savez(lpath, labels=np.array(labels, dtype=object))
    return cpath, lpath


@pytest.mark.parametrize("kind,mix", [("aig", "aig"), ("xmg", "xmg")])
def test_npz_parser_both_layouts(tmp_path, kind, mix):
    import deepgate
    from deepgate import synth
    circuits = synth.make_circuits(mix, 10, 6, 40, cfg=60, window=12, n_pairs=8)
    root = str(tmp_path / (kind + "_npz"))
    cpath, lpath = write_npz_dataset(root, kind, circuits, extra_names=("D_FF_0", "dlatch", "nopairs_a"))
    torch.manual_seed(3)
    parser = deepgate.NpzParser(root, cpath, lpath, kind)
    train, val = parser.get_dataset()
    # 13 stored; 2 dropped by name; the pair-less one is dropped through ``len(tt_pair_index) == 0`` (parser.py:109) --
    # which an AIG-layout [2, 0] array does not trigger (its len is 2), exactly as upstream; 90 / 10 split
    kept = 11 if kind == "aig" else 10
    assert len(train) == int(kept * 0.9) and len(val) == kept - int(kept * 0.9)
    assert os.path.exists(os.path.join(root, "inmemory", "data.pt"))           # cached like parser.py:55-66
    names = sorted(g.name for g in list(train.graphs) + list(val.graphs))
    assert names == ["c%03d" % i for i in range(10)] + (["nopairs_a"] if kind == "aig" else [])
    by_name = {g.name: g for g in list(train.graphs) + list(val.graphs)}
    for i, c in enumerate(circuits):
        g = by_name["c%03d" % i]
        want = deepgate.parse_pyg_mlpgate(c["x"], c["edge_index"], c["prob"], c["tt_sim"], c["tt_pair_index"])
        for key in ("x", "edge_index", "tt_pair_index", "tt_sim", "forward_level", "backward_level", "forward_index", "prob"):
            assert torch.equal(g[key], want[key]), key
        assert torch.equal(g.gate.reshape(-1).float(), want.gate.reshape(-1))
    # second construction reads the cache (no re-parse) and, unshuffled, keeps file order
    again = deepgate.NpzParser(root, cpath, lpath, kind, random_shuffle=False)
    assert [g.name for g in again.train_dataset.graphs] == ["c%03d" % i for i in range(9)]
    assert len(again.train_dataset) + len(again.val_dataset) == kept
    # the loader the Trainer builds (trainer.py:189-195)
    loader = deepgate.DataLoader(train, batch_size=4, shuffle=False, drop_last=True)
    batches = list(loader)
    assert len(batches) == 2 and all(b.num_graphs == 4 for b in batches)
    assert int(batches[0].edge_index.max()) < batches[0].x.size(0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.gpu
@pytest.mark.parametrize("kind,mix,model_name", [("mig", "mig4", "DG_AE"), ("aig", "aig", "DG_AE"), ("xmg", "xmg", "DG_VAE")])
def test_replay_of_train_py_call_sequence(tmp_path, monkeypatch, kind, mix, model_name):
    """train.py:23-104: NpzParser -> DirectMultiGCNEncoder -> Model -> Trainer(args, model, training_id=, batch_size=,
    distributed=True) -> [resume] -> per stage set_training_args + train + save; then a second Trainer resumes."""
    import deepgate
    from deepgate import synth
    circuits = synth.make_circuits(mix, 12, 8, 60, cfg=61, window=16, n_pairs=12)
    root = str(tmp_path / "data")
    cpath, lpath = write_npz_dataset(root, kind, circuits)
    args = types.SimpleNamespace(model=model_name, type=kind, dim_hidden=64, dim_feature=6, s_rounds=2, t_rounds=2,
                                 layernorm=True, batch_size=4, exp_id="replay_%s" % kind, resume=False)
    for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", str(_free_port())), ("RANK", "0"), ("WORLD_SIZE", "1"),
                 ("LOCAL_RANK", "0")):
        monkeypatch.setenv(k, v)
    monkeypatch.chdir(tmp_path)                                               # Trainer's default save_dir is ./exp
    torch.manual_seed(0)
    dataset = deepgate.NpzParser(root, cpath, lpath, args.type)
    train_dataset, val_dataset = dataset.get_dataset()
    model_map = {"aig": deepgate.dg_ae_model_aig.Model, "mig": deepgate.dg_ae_model_mig.Model,
                 "xmg": deepgate.dg_ae_model_xmg.Model, "xag": deepgate.dg_ae_model_xag.Model}

    def make_model():
        encoder = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=args.dim_hidden, dim_feature=args.dim_feature,
                                                             enable_reverse=True, s_rounds=args.s_rounds,
                                                             t_rounds=args.t_rounds, layernorm=args.layernorm)
        kw = {"variational": True} if "VAE" in args.model else {}
        return model_map[args.type](struct_encoder=encoder, dim_hidden=args.dim_hidden, enable_encode=True,
                                    enable_reverse=True, **kw)

    try:
        trainer = deepgate.Trainer(args, make_model(), training_id=args.exp_id, batch_size=args.batch_size, distributed=True)
        assert trainer.resume() is False
        before = {k: v.detach().clone() for k, v in trainer.model.state_dict().items()}
        for stage, weights in enumerate(([1.0, 0.0, 0.0], [1.0, 4.0, 4.0])):
            trainer.set_training_args(rc_prob_func_weight=weights, lr=1e-3, lr_step=50)
            trainer.train(1, train_dataset, val_dataset)
            if trainer.local_rank == 0:
                trainer.save(os.path.join(trainer.log_dir, "stage_%d.pth" % (stage + 1)))
        after = trainer.model.state_dict()
        # the steps trained the model: every tensor moved, except those whose gradient is mathematically zero (the query
        # path / key bias / attention bias cancel in the softmax; fan-in-1 gates have no attention gradient at all; in the
        # single-round sweep -- num_rounds = 1, the reference default -- the GRU runs from h = 0, so d weight_hh = d gh (x) h = 0)
        frozen = (".msg_q.", "msg_k.bias", "attn_lin.bias", "aggr_not_func.attn_lin", "aggr_not_func.msg_k", "_func.weight_hh_l0")
        still = [k for k in before if before[k].is_floating_point() and torch.equal(before[k], after[k])
                 and not any(tag in k for tag in frozen)]
        assert not still, still
        assert all(torch.isfinite(v).all() for v in after.values() if v.is_floating_point())
        log = open(trainer.log_path).read()
        assert log.count("train|") == 2 and log.count("val|") == 2
        assert os.path.exists(os.path.join(trainer.log_dir, "model_last.pth"))
        if "VAE" in args.model:                                                # trainer.py:145-151 branch
            batch = next(iter(deepgate.DataLoader(train_dataset, batch_size=4))).to(trainer.device)
            status = trainer.run_batch(batch)
            assert "kl_loss" in status and float(status["kl_loss"]) > 0.0
        # a fresh process would do: Trainer(...).resume() -> continues from the saved epoch with the saved weights
        other = deepgate.Trainer(args, make_model(), training_id=args.exp_id, batch_size=args.batch_size, distributed=True)
        assert other.resume() is True and other.model_epoch == 0               # model_last.pth is written at epoch 0 (trainer.py:262-264)
        saved = torch.load(os.path.join(trainer.log_dir, "model_last.pth"), weights_only=False)["state_dict"]
        for k, v in other.model.state_dict().items():
            assert torch.equal(v.cpu(), saved[k].cpu()), k
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
