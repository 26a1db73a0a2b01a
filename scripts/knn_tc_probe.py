"""Times the batched cosine top-k (tensor-core shortlist path vs the exact streaming path): python scripts/knn_tc_probe.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dcnr_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    only_q = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for d in ((16,) if only_q else (16, 64)):
        E = torch.randn(n, d, device=dev, generator=g)
        model = dcnr_b200.NearestNeighbors().fit(E)
        for nq in ((only_q,) if only_q else (8, 32, 256, 1024)):
            Q = E[torch.randint(0, n, (nq,), device=dev, generator=g)] + 0.05 * torch.randn(nq, d, device=dev, generator=g)
            res = {}
            for name, mn in (("tc", 8), ("stream", 1 << 30)):
                if name == "stream" and nq > 256 and d == 64:
                    continue
                model.tc_min_queries = mn
                reps = 5 if (name == "tc" or nq <= 32) else 2
                secs = bench.time_steps(lambda: model.kneighbors_tensor(Q, 201), reps, 1, lambda: None) / reps
                res[name] = (secs, model.kneighbors_tensor(Q, 201))
                print(f"n {n} d {d} Q {nq:5d} {name:7s}: {secs * 1e3:9.3f} ms  {n * nq / secs / 1e9:9.1f} G pairs/s", flush=True)
            if len(res) == 2:
                same = torch.equal(res["tc"][1][1], res["stream"][1][1]) and torch.equal(res["tc"][1][0], res["stream"][1][0])
                print(f"    bit-exact tc vs stream: {same}", flush=True)
        del E, model


if __name__ == "__main__":
    main()
