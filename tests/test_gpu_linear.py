"""csrc/linear.cu (weight / bias gradient of the per-node nn.Linear layers) and csrc/linear_tc.cu (their forward and data
gradient on the tensor cores, 64 / 128-wide layers) against torch.autograd in float64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,I,O,bias", [(1000, 128, 64, True), (4097, 64, 128, True), (777, 64, 32, True), (5000, 32, 32, True),
                                        (3001, 32, 1, True), (33, 70, 17, False), (1, 6, 3, True),
                                        # tensor-core shapes: one row, a ragged last tile, more tiles than SMs, no bias
                                        (1, 64, 64, True), (129, 128, 128, True), (65818, 128, 64, True), (40000, 64, 128, False),
                                        (300, 64, 64, False)])
def test_linear_matches_autograd(N, I, O, bias):
    from deepgate import ops
    g = torch.Generator().manual_seed(N + I + O)
    x = torch.randn(N, I, generator=g).cuda().requires_grad_(True)
    lin = ops.Linear(I, O, bias=bias).cuda()
    gy = torch.randn(N, O, generator=g).cuda()
    y = lin(x)
    y.backward(gy)
    xd = x.detach().double().requires_grad_(True)
    wd = lin.weight.detach().double().requires_grad_(True)
    bd = lin.bias.detach().double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(gy.double())
    tol = 1e-5
    assert float((y.double() - yd).abs().max()) <= tol * max(1.0, float(yd.abs().max()))
    assert float((x.grad.double() - xd.grad).abs().max()) <= tol * max(1.0, float(xd.grad.abs().max()))
    assert float((lin.weight.grad.double() - wd.grad).abs().max()) <= tol * max(1.0, float(wd.grad.abs().max()))
    if bias:
        assert float((lin.bias.grad.double() - bd.grad).abs().max()) <= tol * max(1.0, float(bd.grad.abs().max()))


def test_tensor_core_path_is_taken_and_matches_the_library_path(monkeypatch):
    from deepgate import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, 128, generator=g).cuda()
    lin = ops.Linear(128, 64).cuda()
    ops.PROFILE = {}
    y = lin(x)
    torch.cuda.synchronize()
    assert "linear_tc" in ops.PROFILE
    ops.PROFILE = None
    monkeypatch.setenv("MGV_LINEAR_TORCH", "1")
    y_lib = lin(x)
    assert float((y - y_lib).abs().max()) <= 2e-6 * float(y_lib.abs().max())


def test_linear_state_dict_keys_are_nn_linear():
    from deepgate import ops
    assert set(ops.Linear(8, 4).state_dict()) == set(torch.nn.Linear(8, 4).state_dict())
