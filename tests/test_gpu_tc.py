"""tcgen05 building blocks (csrc/mgv_tc.cuh) against float64 matmuls: every operand orientation the
fused kernels use, through the library's diagnostic entry point mgv_tc_selftest."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, A, B, K, N, nout):
    from deepgate import _native as nat
    D = torch.full((128, nout), float("nan"), dtype=torch.float32, device="cuda")
    nat.check(nat.lib().mgv_tc_selftest(mode, nat.ptr(A), nat.ptr(B), nat.ptr(D), K, N, nat.stream_of(A.device)),
              "mgv_tc_selftest")
    torch.cuda.synchronize()
    return D


def _rel(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("K,N", [(64, 64), (128, 192), (128, 256), (64, 16)])
def test_k_major_sw128(K, N):
    g = torch.Generator().manual_seed(K + N)
    A = torch.randn(128, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) * 0.125).cuda()
    assert _rel(_run(0, A, B, K, N, N), A.double() @ B.double().t()) < 2e-6


@pytest.mark.parametrize("N", [64, 256])
def test_mn_major_both(N):
    g = torch.Generator().manual_seed(N)
    X = torch.randn(128, 128, generator=g).cuda()
    Y = torch.randn(128, N, generator=g).cuda()
    assert _rel(_run(1, X, Y, 0, N, N), X.double().t() @ Y.double()) < 2e-6


def test_plain16_k_major():
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 16, generator=g).cuda()
    B = torch.randn(256, 16, generator=g).cuda()
    assert _rel(_run(2, A, B, 16, 256, 256), A.double() @ B.double().t()) < 2e-6


def test_mn_sw128_times_mn_plain16():
    g = torch.Generator().manual_seed(6)
    X = torch.randn(128, 128, generator=g).cuda()
    Z = torch.randn(128, 16, generator=g).cuda()
    assert _rel(_run(3, X, Z, 0, 16, 16), X.double().t() @ Z.double()) < 2e-6


def test_k_major_times_mn_major_weights():
    g = torch.Generator().manual_seed(7)
    A = torch.randn(128, 192, generator=g).cuda()
    W = (torch.randn(192, 64, generator=g) * 0.125).cuda()
    assert _rel(_run(4, A, W, 192, 64, 64), A.double() @ W.double()) < 2e-6


def test_small_magnitudes_and_saturation():
    """fp16 hi/lo split: absolute error floor 2^-25 per operand element; |x| > 65504 saturates (finite)."""
    g = torch.Generator().manual_seed(8)
    A = (torch.randn(128, 64, generator=g) * 1e-3).cuda()
    B = torch.randn(64, 64, generator=g).cuda()
    ref = A.double() @ B.double().t()
    assert float((_run(0, A, B, 64, 64, 64).double() - ref).abs().max()) < 64 * 8 * 2.0 ** -24
    A2 = A.clone()
    A2[0, 0] = 1e6
    assert torch.isfinite(_run(0, A2, B, 64, 64, 64)).all()


def test_a_operand_from_tensor_memory():
    """dG (fp16 hi/lo written to tensor memory by tcgen05.st) x MN-major weights, N = 128 from two 64-column blocks:
    the data-gradient product of the struct-encoder backward."""
    g = torch.Generator().manual_seed(9)
    A = torch.randn(128, 64, generator=g).cuda()
    W = (torch.randn(64, 128, generator=g) * 0.125).cuda()
    assert _rel(_run(5, A, W, 64, 128, 128), A.double() @ W.double()) < 2e-6
