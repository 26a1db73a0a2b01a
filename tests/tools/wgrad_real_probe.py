"""Weight-gradient error of the tensor-core kernel on the REAL operands of a training step (dz and the layer input X captured from
the reference class run in float64 on the GPU), plain and with X centred per column:
    dW = dz^T X = dz^T (X - mean) + colsum(dz) (x) mean
Usage: python tests/tools/wgrad_real_probe.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dcnr_b200  # noqa: E402,F401
from oracle import dcnr_oracle as orc  # noqa: E402
from scripts.wgrad_center_probe import err, wgrad  # noqa: E402
from tests.helpers import synth_inputs  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    ref_main, _ = bench.load_reference()
    nu, ni = 200_000, 50_000
    params = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0)
    state = orc.make_state(nu, ni, bench.CAT_DIMS, bench.N_NUM, params, seed=3, emb_scale=0.1, randomize_bn=True)
    m = ref_main.DCN_RecSys(nu, ni, bench.CAT_DIMS, bench.N_NUM, params)
    m.load_state_dict(state)
    m = m.double().cuda().train()
    u, i, c, x, y = synth_inputs(nu, ni, bench.CAT_DIMS, bench.N_NUM, B, seed=7, zipf=True, device="cuda")
    cap = {}

    def hook(name):
        def f(mod, inp, out):
            out.retain_grad()
            cap[name] = (inp[0].detach(), out)
        return f
    layers = {"initial_deep_layer": m.initial_deep_layer}
    for r, blk in enumerate(m.res_blocks):
        layers[f"res_blocks.{r}.layer1"] = blk.layer1
        layers[f"res_blocks.{r}.layer2"] = blk.layer2
    for n, l in layers.items():
        l.register_forward_hook(hook(n))
    logits = m(u, i, c, x.double())
    loss = torch.nn.BCEWithLogitsLoss()(logits, y.double())
    loss.backward()
    for n in layers:
        X64, out = cap[n]
        dz64 = out.grad
        ref = dz64.t() @ X64
        k = X64.shape[1]
        kp = (k + 31) // 32 * 32
        X = torch.zeros(B, kp, device="cuda")
        X[:, :k] = X64.float()
        dz = dz64.float().contiguous()
        mu = X.double().mean(0)
        Xc = (X.double() - mu).float()
        corr = torch.outer(dz.double().sum(0), mu)
        ratio = float((X64.mean(0).abs() / X64.std(0).clamp_min(1e-30)).max())
        line = [f"{n:28s} max|mean|/std {ratio:7.2f}"]
        for prec in ("fp32", "tf32x3"):
            line.append(f"{prec} plain {err(wgrad(dz, X, prec)[:, :k], ref):.2e}")
            line.append(f"centred {err((wgrad(dz, Xc, prec).double() + corr)[:, :k], ref):.2e}")
        print("  ".join(line), flush=True)


if __name__ == "__main__":
    main()
