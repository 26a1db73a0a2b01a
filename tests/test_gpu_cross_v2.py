"""GPU parity of the opt-in DCN-v2 cross network (dcnr_cross_v2_fwd / dcnr_cross_v2_bwd_prep + the linear entry points)
against oracle/cross_v2_oracle.py.  The oracle is not pinned to the reference (the reference has no such layer)."""
import numpy as np
import pytest
import torch

import dcnr_b200
from oracle import cross_v2_oracle as V2

pytestmark = pytest.mark.gpu


def _case(B, D, L, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, D, generator=g) * 0.5
    ws = [torch.randn(D, D, generator=g) / D ** 0.5 for _ in range(L)]
    bs = [torch.randn(D, generator=g) * 0.1 for _ in range(L)]
    gy = torch.randn(B, D, generator=g)
    return x0, ws, bs, gy


def _nerr(a, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


# tolerance: max-abs-normalised error, the metric SURVEY 8d uses for logits and gradients.  fp32 / tf32x3 must meet the
# 1e-5 parity bar; tf32 (one truncating MMA) is a stated tolerance.
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32x3", 1e-5), ("tf32", 2e-2)])
@pytest.mark.parametrize("B,D,L", [(4096, 57, 3), (1000, 64, 1), (333, 175, 2), (1, 57, 2)])
def test_cross_v2_forward_backward_match_oracle(precision, tol, B, D, L):
    x0, ws, bs, gy = _case(B, D, L, 11 + B + D)
    dev = torch.device("cuda")
    net = dcnr_b200.CrossNetworkV2(D, L, precision).to(dev)
    with torch.no_grad():
        for l, layer in enumerate(net.layers):
            layer.w.weight.copy_(ws[l])
            layer.b.copy_(bs[l])
    xg = x0.to(dev).requires_grad_()
    y = net(xg)
    assert y.shape == (B, D) and y.dtype == torch.float32
    y.backward(gy.to(dev))
    y_ref, dx0_ref, gws_ref, gbs_ref = V2.cross_v2_numpy(x0.numpy(), [w.numpy() for w in ws], [b.numpy() for b in bs], gy.numpy())
    assert _nerr(y.detach().cpu().numpy(), y_ref) <= tol
    assert _nerr(xg.grad.cpu().numpy(), dx0_ref) <= tol
    for l, layer in enumerate(net.layers):
        assert _nerr(layer.w.weight.grad.cpu().numpy(), gws_ref[l]) <= tol, f"dW[{l}]"
        assert _nerr(layer.b.grad.cpu().numpy(), gbs_ref[l]) <= tol, f"db[{l}]"


def test_cross_v2_large_batch_uses_the_tensor_core_kernel_and_is_linear_in_the_upstream_gradient():
    """Size-independent property at a bench-sized batch: the backward is linear in gy, and the forward of a zero-weight
    layer is the identity plus x0 * b."""
    dev = torch.device("cuda")
    B, D = 1 << 18, 57
    g = torch.Generator(device=dev).manual_seed(3)
    x0 = torch.randn(B, D, generator=g, device=dev) * 0.5
    net = dcnr_b200.CrossNetworkV2(D, 2, "tf32x3").to(dev)
    n0 = dcnr_b200.launch_count()
    with torch.no_grad():
        y = net(x0)
    assert dcnr_b200.launch_count() > n0
    with torch.no_grad():
        for layer in net.layers:
            layer.w.weight.zero_()
            layer.b.fill_(0.25)
        y0 = net(x0)
    # x1 = x0 * 0.25 + x0 ; x2 = x0 * 0.25 + x1 = 1.5 x0
    torch.testing.assert_close(y0, 1.5 * x0, rtol=1e-6, atol=1e-6)
    assert torch.isfinite(y).all()


def test_cross_layer_v2_rejects_cpu_tensors():
    layer = dcnr_b200.CrossLayerV2(8)
    with pytest.raises(RuntimeError):
        layer(torch.randn(4, 8))
