"""Operator-level parity of the dense-layer GEMMs through the C ABI (pytest -m gpu):
the CUDA-core fp32 path, the tcgen05 3xTF32 parity path and the single-pass TF32 path, against a
float64 matmul.  Tolerances (max-abs error / max-abs reference): fp32 and tf32x3 <= 2e-6 (the
contract's 1e-5 leaves room for five chained layers), tf32 <= 2e-3 (stated fast-path tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-6, "tf32x3": 6e-6, "tf32": 2e-3}
SHAPES = [(1000, 256, 256), (128, 256, 64), (4133, 512, 256), (300, 96, 96), (5000, 32, 32), (77, 256, 64),
          (2048, 160, 224)]


def _err(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_linear_forward_epilogues(precision, m, n, k):
    import dcnr_b200
    F_ = dcnr_b200.functional
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    s = torch.rand(n, device="cuda", generator=g) + 0.5
    r = torch.randn(m, n, device="cuda", generator=g)
    acc = x.double() @ w.double().t()
    y = F_.linear_forward_raw(x, w, b, precision=precision)
    assert _err(y, acc + b.double()) < TOL[precision]
    y = F_.linear_forward_raw(x, w, b, s, r, True, precision)
    assert _err(y, torch.relu(acc * s.double() + b.double() + r.double())) < TOL[precision]
    y = F_.linear_forward_raw(x, w, None, None, None, False, precision)
    assert _err(y, acc) < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("m,n,k", [(1000, 256, 256), (4133, 256, 64), (300, 96, 96), (513, 512, 512), (70001, 256, 256),
                                   (65536, 128, 32), (31, 384, 224)])
def test_linear_autograd(precision, m, n, k):
    import dcnr_b200
    F_ = dcnr_b200.functional
    g = torch.Generator(device="cuda").manual_seed(7 * m + n + k)
    x = torch.randn(m, k, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, device="cuda", generator=g, requires_grad=True)
    gy = torch.randn(m, n, device="cuda", generator=g)
    y = F_.linear(x, w, b, precision)
    y.backward(gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    (xd @ wd.t() + bd).backward(gy.double())
    tol = TOL[precision]
    assert _err(x.grad, xd.grad) < tol
    assert _err(w.grad, wd.grad) < max(tol, 2e-6)      # tcgen05 MN-major wgrad when n % 128 == 0 and k % 32 == 0 (k <= 256)
    assert _err(b.grad, bd.grad) < 2e-6


def test_unaligned_leading_dimension_uses_cuda_core_path():
    import dcnr_b200
    F_ = dcnr_b200.functional
    x = torch.randn(333, 57, device="cuda")
    w = torch.randn(256, 57, device="cuda") / 8
    for precision in ("fp32", "tf32x3"):
        y = F_.linear_forward_raw(x, w, None, precision=precision)
        assert _err(y, x.double() @ w.double().t()) < 2e-6
