"""How does the tcgen05 fp32 accumulator round?  (GPU probe, round 2.)

Feeds the library's dense-layer GEMM (dcnr_linear_fwd) operands whose products are EXACT in fp32
(tf32-representable inputs, single-pass tf32 mode), so the only error left is the accumulation inside
the tensor core, and reports the SIGNED mean error against float64 for all-positive, all-negative and
mixed-sign sums.  Round-to-nearest gives a mean of ~0 in all three; truncation toward zero gives
- / + / ~0; truncation toward -inf (two's complement) gives - / - / -.
Then the same statistics for the 3-term split (tf32x3) and the CUDA-core fp32 path on general inputs,
including the error of COLUMN SUMS of the output (what a BatchNorm bias gradient sees).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcnr_b200  # noqa: E402
from dcnr_b200 import functional as F  # noqa: E402


def to_tf32(t):
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def stats(name, c, ref):
    err = c.double() - ref
    scale = ref.abs().mean()
    eps = 2.0 ** -24
    print(f"{name:44s} mean signed err / (eps*mean|C|) = {float(err.mean() / (eps * scale)):+9.3f}   rms = "
          f"{float(err.pow(2).mean().sqrt() / (eps * scale)):8.3f}   colsum err / max|colsum| = "
          f"{float((err.sum(0)).abs().max() / ref.sum(0).abs().max()):.2e}")


def main():
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    M, N = 16384, 256
    for K in (64, 256):
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / 16
        at, wt = to_tf32(a), to_tf32(w)
        print(f"--- K = {K}: single tf32 MMA pass on tf32-exact operands (pure accumulation error)")
        for tag, aa, ww in (("all products > 0", at.abs(), wt.abs()), ("all products < 0", at.abs(), -wt.abs()),
                            ("mixed signs", at, wt)):
            ref = aa.double() @ ww.double().t()
            c = F.linear_forward_raw(aa.to(dev), ww.to(dev), None, None, None, False, "tf32").cpu()
            stats(f"tf32 exact operands, {tag}", c, ref)
            c = F.linear_forward_raw(aa.to(dev), ww.to(dev), None, None, None, False, "fp32").cpu()
            stats(f"fp32 CUDA cores,     {tag}", c, ref)
        print(f"--- K = {K}: general fp32 operands")
        ref = a.double() @ w.double().t()
        for prec in ("tf32x3", "fp32"):
            c = F.linear_forward_raw(a.to(dev), w.to(dev), None, None, None, False, prec).cpu()
            stats(f"{prec} mixed signs", c, ref)
        ref = a.abs().double() @ w.abs().double().t()
        for prec in ("tf32x3", "fp32"):
            c = F.linear_forward_raw(a.abs().to(dev), w.abs().to(dev), None, None, None, False, prec).cpu()
            stats(f"{prec} all products > 0", c, ref)
        ref = torch.matmul(a.double(), w.double().t())
        c = (a.to(dev) @ w.to(dev).t()).cpu()
        stats("torch.matmul fp32 (cuBLAS, allow_tf32 off)", c, ref)


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    main()
