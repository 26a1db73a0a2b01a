"""Times the fused eval tower alone (dcnr_tower_eval on a resident x0) and the whole eval forward per precision.
Usage: python scripts/tower_probe.py [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dcnr_b200  # noqa: E402
from dcnr_b200 import _cabi as C  # noqa: E402

FLOP_TOWER = 2 * 57 * 256 + 4 * 2 * 256 * 256


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    dev = torch.device("cuda")
    m = bench.synth_state_device(dev).eval()
    dims, ps = m._dims(), m._param_struct()
    x0 = torch.zeros(rows, dims.in_dim_pad, device=dev)
    x0[:, :57] = torch.randn(rows, 57, device=dev) * 0.3
    cross = torch.randn(rows, device=dev)
    out = torch.empty(rows, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = torch.empty(C.lib().dcnr_tower_eval_workspace_bytes(dims), dtype=torch.uint8, device=dev)
    for prec in ("fp16x3", "bf16"):
        for options, tag in ((0, "one CTA per SM"),):
            def run():
                C.check(C.lib().dcnr_tower_eval(dims, ps, C.ptr(x0), x0.shape[1], C.ptr(cross), C.ptr(out), rows, C.PRECISIONS[prec],
                                                options, C.ptr(flags), C.ptr(ws), ws.numel(), C.stream()))
            secs = bench.time_steps(run, 10, 3, lambda: None) / 10
            print(f"tower {prec:7s} {tag:12s} rows {rows}: {secs * 1e3:8.3f} ms  {rows / secs / 1e6:8.1f} M rows/s  "
                  f"{FLOP_TOWER * rows / secs / 1e12:7.1f} TFLOP/s algorithmic  flags {int(flags[0].item())}", flush=True)
    u, i, c, x = bench.synth_requests(max(1, rows // 500), 500, 1234, dev)
    for prec in ("fp16x3", "bf16", "tf32x3"):
        m.precision = prec
        with torch.no_grad():
            secs = bench.time_steps(lambda: m(u, i, c, x), 5, 2, lambda: None) / 5
        print(f"eval forward {prec:7s} rows {u.numel()}: {secs * 1e3:8.3f} ms  {u.numel() / secs / 1e6:8.1f} M candidates/s", flush=True)


if __name__ == "__main__":
    main()
