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
