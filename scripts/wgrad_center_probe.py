"""Does centring the X operand help the tensor-core weight gradient?  dW = dz^T X with sum_b dz = 0 per column (BatchNorm backward)
and X with non-zero column means: the mean part cancels exactly in theory, its rounding error does not.
Usage: python scripts/wgrad_center_probe.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcnr_b200  # noqa: E402,F401
from dcnr_b200 import _cabi as C  # noqa: E402


def wgrad(dz, X, prec):
    m, n = dz.shape
    k = X.shape[1]
    dw = torch.empty(n, k, device="cuda")
    sb = C.lib().dcnr_linear_wgrad_scratch_bytes(m, n, k)
    scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
    C.check(C.lib().dcnr_linear_wgrad(C.ptr(dz), dz.stride(0), C.ptr(X), X.stride(0), C.ptr(dw), k, None, m, n, k,
                                      C.PRECISIONS[prec], C.ptr(scratch), sb, C.stream()))
    torch.cuda.synchronize()
    return dw


def err(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    g = torch.Generator(device="cuda").manual_seed(5)
    for mean_scale in (0.0, 0.5, 2.0):
        mu = torch.randn(256, generator=g, device="cuda") * mean_scale
        X = torch.randn(B, 256, generator=g, device="cuda") + mu
        dz = torch.randn(B, 256, generator=g, device="cuda") * 1e-5
        dz = dz - dz.mean(0, keepdim=True)
        ref = dz.double().t() @ X.double()
        colmean = X.double().mean(0)
        Xc = (X.double() - colmean).float()
        corr = torch.outer(dz.double().sum(0), colmean)
        line = [f"B {B} mean scale {mean_scale}:"]
        for prec in ("fp32", "tf32x3"):
            line.append(f"{prec} plain {err(wgrad(dz, X, prec), ref):.2e}")
            line.append(f"{prec} centred {err(wgrad(dz, Xc, prec).double() + corr, ref):.2e}")
        print("  ".join(line), flush=True)


if __name__ == "__main__":
    main()
