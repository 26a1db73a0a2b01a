"""GPU parity tests of the cosine top-k path through the C ABI (pytest -m gpu).

Bit-exact contract: distances and indices equal to oracle/knn_oracle.c (sequential-fma fp32,
order (dist asc, idx asc)); equal to scikit-learn's golden outputs on tie-free data.
"""
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small", "k51", "d64"])
def test_matches_sklearn_golden(name):
    import dcnr_b200
    z = np.load(os.path.join(GOLDEN, f"knn_{name}.npz"))
    nn_model = dcnr_b200.NearestNeighbors(n_neighbors=16, metric="cosine", algorithm="brute")     # main.py:268
    nn_model.fit(z["E"])                                                                            # main.py:269
    k = int(z["k"])
    for q in range(z["Q"].shape[0]):
        dist, ind = nn_model.kneighbors(z["Q"][q].reshape(1, -1), n_neighbors=k)                    # main.py:200,300
        assert dist.dtype == np.float32 and ind.dtype == np.int64 and dist.shape == (1, k)
        assert np.abs(dist[0] - z["dist"][q]).max() < 5e-7
        gaps = np.diff(z["dist"][q])
        safe = np.r_[True, gaps > 1e-6] & np.r_[gaps > 1e-6, True]
        assert (ind[0][safe] == z["ind"][q][safe]).all()
        assert ind[0][0] == z["q_rows"][q]


@pytest.mark.parametrize("n,d,nq,k", [(100_000, 16, 3, 11), (100_000, 16, 9, 201), (50_000, 32, 2, 51),
                                      (30_000, 24, 2, 16), (1000, 16, 1, 256), (257, 64, 17, 5)])
def test_bit_exact_against_oracle(n, d, nq, k):
    import dcnr_b200
    rng = np.random.default_rng(n + d + k)
    E = rng.standard_normal((n, d)).astype(np.float32)
    Q = E[rng.integers(0, n, nq)] + (0.01 * rng.standard_normal((nq, d))).astype(np.float32)
    ref_d, ref_i = knn_oracle.OracleNearestNeighbors().fit(E).kneighbors(Q, n_neighbors=k)
    got_d, got_i = dcnr_b200.NearestNeighbors().fit(E).kneighbors(Q, n_neighbors=k)
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_d.view(np.uint32), ref_d.view(np.uint32))


def test_ties_broken_by_index_and_zero_rows():
    import dcnr_b200
    rng = np.random.default_rng(1)
    E = rng.standard_normal((5000, 16)).astype(np.float32)
    E[4000] = E[123]; E[77] = E[123]; E[10] = 0.0          # exact duplicates and a zero-norm row
    model = dcnr_b200.NearestNeighbors().fit(E)
    d, i = model.kneighbors(E[123].reshape(1, -1), n_neighbors=5)
    assert list(i[0][:3]) == [77, 123, 4000]
    ref = knn_oracle.OracleNearestNeighbors().fit(E).kneighbors(E[[123, 10]], n_neighbors=20)
    got = model.kneighbors(E[[123, 10]], n_neighbors=20)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
    with pytest.raises(ValueError):
        model.kneighbors(E[:1], n_neighbors=6000)


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_merge_is_shard_count_independent(shards):
    import dcnr_b200
    rng = np.random.default_rng(2)
    E = rng.standard_normal((20_000, 16)).astype(np.float32)
    E[15000] = E[5]
    Q = E[[5, 900, 19999]]
    full_d, full_i = dcnr_b200.NearestNeighbors().fit(E).kneighbors(Q, n_neighbors=201)
    bounds = np.linspace(0, E.shape[0], shards + 1).astype(int)
    dparts, iparts = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        m = dcnr_b200.NearestNeighbors(index_base=int(a)); m._allow_short = True
        m.fit(E[a:b])
        dd, ii = m.kneighbors_tensor(torch.from_numpy(Q).cuda(), 201)
        dparts.append(dd); iparts.append(ii)
    md, mi = dcnr_b200.merge_shards(torch.stack(dparts), torch.stack(iparts))
    assert np.array_equal(mi.cpu().numpy(), full_i) and np.array_equal(md.cpu().numpy(), full_d)


def test_full_size_properties():
    """cfg4-sized shard on one GPU (10 M x 16 would be 640 MB: use 4 M to bound test time):
    sorted output, the query row itself first, distances reproduce from the catalog."""
    import dcnr_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    E = torch.randn(4_000_000, 16, device="cuda", generator=g)
    model = dcnr_b200.NearestNeighbors().fit(E)
    rows = torch.tensor([0, 1_234_567, 3_999_999], device="cuda")
    d, i = model.kneighbors_tensor(E[rows], 201)
    assert torch.equal(i[:, 0], rows)
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    ehat = model._catalog_hat
    sims = (ehat[i[1]] * ehat[rows[1]]).sum(1)
    assert float((1 - sims - d[1]).abs().max()) < 1e-5
    # a brute-force check of the k-th distance: nothing outside the list is closer
    all_d = (1 - ehat @ ehat[rows[1]]).clamp_(0, 2)
    assert int((all_d < d[1, -1] - 1e-6).sum()) <= 201


# ---- query batches: tensor-core shortlist + exact re-score (csrc/topk_tc.cu) ------------------------------------------------
def _both_paths(E, Q, k, index_base=0):
    """(tensor-core path, exact streaming path) results for the same fitted catalog."""
    import dcnr_b200
    model = dcnr_b200.NearestNeighbors(index_base=index_base).fit(E)
    from dcnr_b200 import _cabi as C
    assert C.lib().dcnr_knn_tc_supported(E.shape[0], E.shape[1], Q.shape[0], k)
    tc = model.kneighbors_tensor(Q, k)
    model.tc_min_queries = 1 << 30
    exact = model.kneighbors_tensor(Q, k)
    return tc, exact


@pytest.mark.parametrize("n,d,nq,k", [(300_000, 16, 8, 11), (300_000, 16, 33, 201), (270_001, 32, 300, 51), (262_144, 64, 40, 256),
                                      (1_000_000, 16, 1024, 201), (500_000, 64, 700, 16), (300_000, 64, 20, 33), (300_000, 32, 64, 100),
                                      (300_000, 32, 1100, 11),             # more queries than one launch holds (512 at d = 32)
                                      (300_000, 24, 50, 21), (280_000, 48, 9, 64)])     # 8- and 16-float K blocks
def test_batched_queries_bit_exact_with_the_streaming_path(n, d, nq, k):
    g = torch.Generator(device="cuda").manual_seed(n + d + nq)
    E = torch.randn(n, d, device="cuda", generator=g)
    E[1000] = E[17].clone(); E[n - 1] = E[17].clone(); E[5] = 0.0                     # ties across tiles, a zero row
    rows = torch.randint(0, n, (nq,), device="cuda", generator=g)
    rows[0] = 17
    Q = E[rows] + 0.05 * torch.randn(nq, d, device="cuda", generator=g)
    Q[0] = E[17]
    if nq > 2:
        Q[1] = 0.0                                                    # a zero query: every distance is 1, order = index order
        Q[2] = torch.randn(d, device="cuda", generator=g)             # a query far from everything
    (td, ti), (ed, ei) = _both_paths(E, Q, k, index_base=7_000)
    assert torch.equal(ti, ei)
    assert torch.equal(td.view(torch.int32), ed.view(torch.int32))
    assert int(ti[0, 0]) == 7_000 + 17 and int(ti[0, 1]) == 7_000 + 1000 and int(ti[0, 2]) == 7_000 + n - 1


def test_batched_queries_against_oracle():
    rng = np.random.default_rng(5)
    n, d, nq, k = 280_000, 16, 12, 31
    E = rng.standard_normal((n, d)).astype(np.float32)
    Q = E[rng.integers(0, n, nq)] + (0.02 * rng.standard_normal((nq, d))).astype(np.float32)
    ref_d, ref_i = knn_oracle.OracleNearestNeighbors().fit(E).kneighbors(Q, n_neighbors=k)
    import dcnr_b200
    model = dcnr_b200.NearestNeighbors().fit(E)
    assert nq >= model.tc_min_queries
    got_d, got_i = model.kneighbors(Q, n_neighbors=k)
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_d.view(np.uint32), ref_d.view(np.uint32))


def test_shortlist_overflow_falls_back_to_the_streaming_path():
    """60 000 copies of one vector: every copy passes the query's threshold, the per-query list overflows, the status word
    sends the batch down the exact path -- the answer is still the first k copies in index order."""
    g = torch.Generator(device="cuda").manual_seed(9)
    n, d, k = 400_000, 16, 21
    E = torch.randn(n, d, device="cuda", generator=g)
    dup = torch.arange(1000, 400_000, 6, device="cuda")[:60_000]
    E[dup] = E[3].clone()
    Q = torch.cat([E[3:4], torch.randn(9, d, device="cuda", generator=g)])
    from dcnr_b200 import _cabi as C
    import dcnr_b200
    model = dcnr_b200.NearestNeighbors().fit(E)
    qhat = torch.nn.functional.normalize(Q)
    ws = torch.empty(C.lib().dcnr_knn_tc_scratch_bytes(n, d, 10, k), dtype=torch.uint8, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    dist = torch.empty(10, k, device="cuda"); ind = torch.empty(10, k, dtype=torch.int64, device="cuda")
    C.check(C.lib().dcnr_knn_topk_tc(C.ptr(model._catalog_hat), n, d, C.ptr(qhat), 10, k, 0, C.ptr(dist), C.ptr(ind), C.ptr(ws),
                                     ws.numel(), C.ptr(status), C.stream()))
    assert int(status.item()) & 1
    td, ti = model.kneighbors_tensor(Q, k)
    expect = torch.cat([torch.tensor([3], device="cuda"), dup[: k - 1]])
    assert torch.equal(ti[0], expect)


def test_kneighbors_inside_cuda_graph_capture():
    """The tensor-core path reads a status word back (a sync), which a capturing stream cannot do: under capture the call
    takes the exact streaming path, and the replayed graph gives the eager answer."""
    import dcnr_b200
    g = torch.Generator(device="cuda").manual_seed(3)
    E = torch.randn(300_000, 16, device="cuda", generator=g)
    Q = E[:32].contiguous()
    model = dcnr_b200.NearestNeighbors().fit(E)
    ed, ei = model.kneighbors_tensor(Q, 21)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        model.tc_min_queries = 1 << 30
        model.kneighbors_tensor(Q, 21)                 # warm the allocator on this stream
        model.tc_min_queries = 8
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            gd, gi = model.kneighbors_tensor(Q, 21)
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(gi, ei) and torch.equal(gd, ed)
