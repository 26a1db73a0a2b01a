"""Times the cosine top-k scan through the public NearestNeighbors mirror: python scripts/knn_probe.py [n] [d] [k]
Reports eager (Python call per query batch) and CUDA-graph replay (GPU time only) numbers."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
k = int(sys.argv[3]) if len(sys.argv) > 3 else 201
qs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 8, 32]
g = torch.Generator(device="cuda").manual_seed(7)
E = torch.randn(n, d, device="cuda", generator=g)
nn_ = dcnr_b200.NearestNeighbors().fit(E)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for nq in qs:
    Q = E[torch.randint(0, n, (nq,), device="cuda", generator=g)].contiguous()
    eager = timed(lambda: nn_.kneighbors_tensor(Q, k))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        nn_.kneighbors_tensor(Q, k)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            out = nn_.kneighbors_tensor(Q, k)
    torch.cuda.synchronize()
    graph = timed(gr.replay)
    passes = 1 if nq == 1 else (nq + 7) // 8
    print(f"n={n} d={d} k={k} q={nq}: eager {eager:.3f} ms, graph {graph:.3f} ms -> one-read {n*d*4/graph/1e6:.0f} GB/s, "
          f"per-pass {passes*n*d*4/graph/1e6:.0f} GB/s")
