"""Fused readout head (csrc/readout.cu) against the reference's module stack (arch/mlp.py:14-56: Linear / BatchNorm1d / ReLU /
Dropout x 2 + Linear), clamp (dg_ae_model_mig.py:150-152) and nn.L1Loss (trainer.py:154-156) run by torch on the same inputs."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multi-gate-vae_b200"), ROOT]
pytestmark = pytest.mark.gpu
TOL = 2e-5


def make(n, seed, p_drop=0.2):
    from deepgate.arch.mlp import MLP
    torch.manual_seed(seed)
    mlp = MLP(64, 32, 1, num_layer=3, p_drop=0.2, norm_layer="batchnorm", act_layer="relu").cuda()
    mlp.fc[3].p = mlp.fc[7].p = p_drop                 # (p_drop = 0 at construction would leave the Dropout modules out)
    with torch.no_grad():
        for bn in (mlp.fc[1], mlp.fc[5]):
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
            bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
        mlp.fc[8].bias.fill_(0.4)                      # predictions on both sides of the clamp
        mlp.fc[8].weight.mul_(3.0)
    x = torch.randn(n, 64, device="cuda")
    target = torch.rand(n, 1, device="cuda")
    return mlp, x, target


def rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-12)


def reference(mlp, x, target, masks=None, p=0.0):
    fc = mlp.fc
    h = fc[2](fc[1](fc[0](x)))
    if masks is not None:
        h = h * masks[0] / (1.0 - p)
    h = fc[6](fc[5](fc[4](h)))
    if masks is not None:
        h = h * masks[1] / (1.0 - p)
    pred = torch.clamp(fc[8](h), min=0.0, max=1.0)
    return pred, torch.nn.L1Loss()(pred, target)


def compare(mlp_a, mlp_b, xa, xb, got, want):
    assert rel(got[0], want[0]) < TOL and abs(float(got[1]) - float(want[1])) < TOL * max(1.0, abs(float(want[1])))
    assert rel(xa.grad, xb.grad) < 5 * TOL
    # scale-relative: the bias of a Linear that feeds a training-mode BatchNorm has a mathematically zero gradient (the
    # normalisation removes the mean) -- both sides hold rounding noise ~1e-7 there
    scale = max(float(pb.grad.abs().max()) for pb in mlp_b.parameters())
    for (k, pa), (_, pb) in zip(mlp_a.named_parameters(), mlp_b.named_parameters()):
        assert float((pa.grad - pb.grad).abs().max()) < 5 * TOL * max(float(pb.grad.abs().max()), 5e-2 * scale), k
    for (k, ba), (_, bb) in zip(mlp_a.named_buffers(), mlp_b.named_buffers()):
        assert rel(ba.float(), bb.float()) < TOL, k


@pytest.mark.parametrize("n", [1, 37, 4099])
@pytest.mark.parametrize("training", [False, True])
def test_fused_head_matches_the_module_stack(n, training):
    """Evaluation mode (running statistics) and training mode without dropout (batch statistics, running-statistics update)."""
    from deepgate import ops
    if training and n == 1:
        pytest.skip("BatchNorm1d needs more than one value per channel in training mode")
    mlp, x, target = make(n, 5, p_drop=0.0 if training else 0.2)
    ref = copy.deepcopy(mlp)
    mlp.train(training); ref.train(training)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    assert mlp.fused_head_ok(xa)
    pred, loss = ops.readout_head(xa, target, mlp)
    (3.0 * loss + (pred * torch.linspace(-1, 1, n, device="cuda").view(n, 1)).sum()).backward()
    rp, rl = reference(ref, xb, target)
    (3.0 * rl + (rp * torch.linspace(-1, 1, n, device="cuda").view(n, 1)).sum()).backward()
    compare(mlp, ref, xa, xb, (pred, loss), (rp, rl))


def test_fused_head_dropout_uses_the_mask_it_reports():
    """Training mode with dropout: the kernel's own keep masks (counter-based hash, exported for the test) applied to the module
    stack give the same prediction, loss and gradients; the keep rate matches 1 - p."""
    from deepgate import ops
    n, p = 3001, 0.2
    mlp, x, target = make(n, 7, p_drop=p)
    ref = copy.deepcopy(mlp)
    mlp.train(); ref.train()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    pred, loss, mask = ops.readout_head(xa, target, mlp, want_mask=True)
    loss.backward()
    bits = ((mask.long().unsqueeze(-1) >> torch.arange(32, device="cuda")) & 1).float()        # [n, 2, 32]
    keep = float(bits.mean())
    assert abs(keep - (1.0 - p)) < 0.01, keep
    rp, rl = reference(ref, xb, target, masks=(bits[:, 0], bits[:, 1]), p=p)
    rl.backward()
    compare(mlp, ref, xa, xb, (pred, loss), (rp, rl))
    # a second call draws a different mask
    _, _, mask2 = ops.readout_head(x, target, mlp, want_mask=True)
    assert not torch.equal(mask, mask2)


def test_model_and_trainer_use_the_fused_head():
    import deepgate
    from deepgate import ops, synth
    from oracle import dg_oracle as O
    G = deepgate.circuits_to_batch(synth.make_circuits("aig", 4, 8, 120, cfg=3), "cuda")
    enc = deepgate.digae_layer.DirectMultiGCNEncoder(dim_hidden=64, dim_feature=6, s_rounds=1, t_rounds=1, layernorm=True)
    model = deepgate.dg_ae_model_aig.Model(struct_encoder=enc, num_rounds=1, dim_hidden=64)
    model.load_state_dict(O.synth_state_dict("aig", 2), strict=False)
    model = model.cuda().eval()
    hs, hf = model(G)
    ops.PROFILE = {}
    pred, loss = model.pred_prob_loss(hf, G.prob)
    torch.cuda.synchronize()
    assert "readout_fwd" in ops.PROFILE
    ops.PROFILE = None
    want = torch.clamp(model.readout_prob(hf), min=0.0, max=1.0)
    assert rel(pred, want) < TOL and abs(float(loss) - float(torch.nn.L1Loss()(want, G.prob))) < TOL
    assert rel(model.pred_prob(hf), want) < TOL
