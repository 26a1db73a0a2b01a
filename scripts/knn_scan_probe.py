"""Times dcnr_knn_topk_tc alone (no fallback): python scripts/knn_scan_probe.py n d Q"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dcnr_b200  # noqa: E402
from dcnr_b200 import _cabi as C  # noqa: E402

n, d, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
k = 201
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
E = torch.randn(n, d, device=dev, generator=g)
model = dcnr_b200.NearestNeighbors().fit(E)
Q = E[torch.randint(0, n, (nq,), device=dev, generator=g)] + 0.05 * torch.randn(nq, d, device=dev, generator=g)
qhat = torch.nn.functional.normalize(Q)
ws = torch.empty(C.lib().dcnr_knn_tc_scratch_bytes(n, d, nq, k), dtype=torch.uint8, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
dist = torch.empty(nq, k, device=dev); ind = torch.empty(nq, k, dtype=torch.int64, device=dev)


def run():
    C.check(C.lib().dcnr_knn_topk_tc(C.ptr(model._catalog_hat), n, d, C.ptr(qhat), nq, k, 0, C.ptr(dist), C.ptr(ind), C.ptr(ws),
                                     ws.numel(), C.ptr(status), C.stream()))


secs = bench.time_steps(run, 5, 2, lambda: None) / 5
print(f"n {n} d {d} Q {nq}: {secs * 1e3:.3f} ms  {n * nq / secs / 1e9:.1f} G pairs/s  status {int(status.item())} KT_DBG={os.environ.get('KT_DBG')}")
