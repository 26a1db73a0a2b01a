"""The oracle against the reference's own outputs (tests/golden, made by make_golden.py).

CPU only.  This is what pins the oracle: every restatement in oracle/ must reproduce the
vectors produced by the unmodified reference classes and scikit-learn.
"""
import os

import numpy as np
import pytest
import torch

from oracle import dcnr_oracle as orc
from oracle import knn_oracle
from tests.helpers import GOLDEN, load_model_case

CASES = ["p0", "p0_trained", "odd", "min"]


@pytest.mark.parametrize("name", CASES)
def test_forward_eval_matches_reference(name):
    c = load_model_case(name)
    out = orc.forward(c["state"], c["user_ids"], c["item_ids"], c["cat"], c["num"], training=False)
    assert out.shape == c["logits_eval"].shape
    assert orc.max_abs_normalised(out, c["logits_eval"]) < 2e-6


@pytest.mark.parametrize("name", CASES)
def test_forward_train_matches_reference(name):
    c = load_model_case(name)
    out = orc.forward(c["state"], c["user_ids"], c["item_ids"], c["cat"], c["num"], training=True)
    assert orc.max_abs_normalised(out, c["logits_train"]) < 5e-6
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in c["state"].items()}
    out64 = orc.forward(st64, c["user_ids"], c["item_ids"], c["cat"], c["num"].double(), training=True)
    assert orc.max_abs_normalised(out64, c["logits_train_f64"]) < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_autograd_backward_matches_reference(name):
    c = load_model_case(name)
    _, grads, _ = orc.forward_backward(c["state"], c["user_ids"], c["item_ids"], c["cat"], c["num"],
                                       grad_logits=c["grad_logits"])
    scale = max(float(g.abs().max()) for g in c["grads"].values())
    for k, ref in c["grads"].items():
        if float(ref.abs().max()) < 1e-6 * scale:          # pre-BN biases: analytically zero
            assert float((grads[k] - ref).abs().max()) < 1e-5 * scale, k
        else:
            assert orc.max_abs_normalised(grads[k], ref) < 2e-4, k   # fp32 vs fp32, different op order


@pytest.mark.parametrize("name", CASES)
def test_numpy_closed_form_matches_reference(name):
    """The hand-derived backward (what the kernels implement) against the reference's autograd."""
    c = load_model_case(name)
    logits, grads = orc.np_forward_backward(c["state"], c["user_ids"], c["item_ids"], c["cat"], c["num"],
                                            c["grad_logits"])
    assert orc.max_abs_normalised(logits, c["logits_train_f64"]) < 1e-10
    # fp64 closed form vs the reference's fp64-free fp32 grads: loose, then tight vs fp64 autograd
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in c["state"].items()}
    _, g64, _ = orc.forward_backward(st64, c["user_ids"], c["item_ids"], c["cat"], c["num"].double(),
                                     grad_logits=c["grad_logits"].double())
    scale = max(float(g.abs().max()) for g in g64.values())
    for k, ref in g64.items():
        got = torch.from_numpy(np.asarray(grads[k])).reshape(ref.shape)
        assert float((got - ref).abs().max()) < 1e-9 * scale, k
    for k, ref in c["grads"].items():
        got = torch.from_numpy(np.asarray(grads[k])).reshape(ref.shape)
        if float(ref.abs().max()) > 1e-6 * scale:
            assert orc.max_abs_normalised(got, ref) < 2e-4, k


@pytest.mark.parametrize("name", ["odd", "min"])
def test_loss_path_matches_reference(name):
    c = load_model_case(name)
    _, grads, loss = orc.forward_backward(c["state"], c["user_ids"], c["item_ids"], c["cat"], c["num"],
                                          labels=c["labels"])
    assert abs(float(loss) - c["loss"]) < 1e-6
    scale = max(float(g.abs().max()) for g in c["lossgrads"].values())
    for k, ref in c["lossgrads"].items():
        assert float((grads[k] - ref).abs().max()) < 2e-4 * scale, k


@pytest.mark.parametrize("name", CASES)
def test_running_stats_update(name):
    c = load_model_case(name)
    st = {k: v.clone() for k, v in c["state"].items()}
    orc.forward(st, c["user_ids"], c["item_ids"], c["cat"], c["num"], training=True, update_running=True)
    for k, ref in c["after"].items():
        if ref.dtype.is_floating_point:
            assert orc.max_abs_normalised(st[k], ref) < 1e-5, k
        else:
            assert int(st[k]) == int(ref), k


def test_cross_layer_closed_form_kat():
    z = np.load(os.path.join(GOLDEN, "cross_layer.npz"))
    x, w, b, g = (z[k].astype(np.float64) for k in ("x", "w", "b", "g"))
    y, _ = orc.np_cross_fwd(x, w[0], b)
    gx, gw, gb = orc.np_cross_bwd(x, w[0], g)
    assert np.abs(y - z["y"]).max() < 1e-5
    assert np.abs(gx - z["gx"]).max() < 1e-4
    assert np.abs(gw - z["gw"][0]).max() < 1e-4
    assert np.abs(gb - z["gb"]).max() < 1e-5
    yt = orc.cross_layer(torch.from_numpy(z["x"]), torch.from_numpy(z["w"]), torch.from_numpy(z["b"]))
    assert np.abs(yt.numpy() - z["y"]).max() < 1e-5


def test_b1_shapes_and_errors():
    c = load_model_case("min")
    out = orc.forward(c["state"], c["user_ids"][:1], c["item_ids"][:1], c["cat"][:1], c["num"][:1], training=False)
    assert out.dim() == 0                                   # .squeeze() -> 0-d at B == 1 (train.py:170)
    with pytest.raises(ValueError):
        orc.forward(c["state"], c["user_ids"][:1], c["item_ids"][:1], c["cat"][:1], c["num"][:1], training=True)


def test_segment_scatter_equals_index_add():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 7, size=200)
    g = rng.standard_normal((200, 5))
    ref = np.zeros((9, 5)); np.add.at(ref, ids, g)
    assert np.abs(orc.segment_scatter(ids, g, 9) - ref).max() < 1e-12


@pytest.mark.parametrize("name", ["small", "k51", "d64"])
def test_knn_oracle_matches_sklearn_golden(name):
    z = np.load(os.path.join(GOLDEN, f"knn_{name}.npz"))
    nn_model = knn_oracle.OracleNearestNeighbors().fit(z["E"])
    k = int(z["k"])
    for i in range(z["Q"].shape[0]):
        dist, ind = nn_model.kneighbors(z["Q"][i].reshape(1, -1), n_neighbors=k)
        assert dist.dtype == np.float32 and ind.dtype == np.int64 and dist.shape == (1, k)
        # tie-free random data: index lists must agree wherever sklearn's distances are separated
        ref_d, ref_i = z["dist"][i], z["ind"][i]
        assert np.abs(dist[0] - ref_d).max() < 5e-7
        gaps = np.diff(ref_d)
        safe = np.r_[True, gaps > 1e-6] & np.r_[gaps > 1e-6, True]
        assert (ind[0][safe] == ref_i[safe]).all()
        assert ind[0][0] == z["q_rows"][i]                  # the query row itself comes first (main.py:201,301 drop it)


def test_knn_oracle_ties_by_index_and_merge():
    rng = np.random.default_rng(3)
    E = rng.standard_normal((500, 16)).astype(np.float32)
    E[400] = E[123]; E[77] = E[123]                         # exact duplicates -> exact distance ties
    nn_model = knn_oracle.OracleNearestNeighbors().fit(E)
    dist, ind = nn_model.kneighbors(E[123].reshape(1, -1), n_neighbors=5)
    assert list(ind[0][:3]) == [77, 123, 400]
    # sharded: per-shard top-k + merge == unsharded, independent of shard count
    ehat = knn_oracle.normalize_rows(E); q = knn_oracle.normalize_rows(E[[123, 5]])
    full = knn_oracle.cosine_topk(ehat, q, 20)
    for shards in (2, 3, 8):
        bounds = np.linspace(0, 500, shards + 1).astype(int)
        parts = [knn_oracle.cosine_topk(ehat[a:b], q, 20, idx_base=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
        d, i = knn_oracle.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
        assert (i == full[1]).all() and (d == full[0]).all()


@pytest.mark.parametrize("name", ["c300", "c40_unmapped", "c12"])
def test_mmr_oracle_matches_reference_rerank(name):
    """oracle/mmr_oracle.c == the reference's rerank_with_mmr (main.py:133-169) on the committed fixtures
    (tests/golden/make_golden.py ran the unmodified function; includes an unmapped best item and C < top_k)."""
    from oracle import mmr_oracle
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"mmr_{name}.npz"))
    got = mmr_oracle.mmr_rerank(z["E"], z["scores"], z["emb_idx"], float(z["lam"]), int(z["top_k"]))
    assert np.array_equal(got, z["order"])


# ------------------------------------------------------------------------------------------------
# opt-in DCN-v2 cross network: the two independent statements of the oracle agree (parity unpinned
# against the reference, which has no such layer -- see oracle/cross_v2_oracle.py)
# ------------------------------------------------------------------------------------------------
def test_cross_v2_oracle_closed_form_matches_autograd():
    from oracle import cross_v2_oracle as V2
    g = torch.Generator().manual_seed(7)
    B, D, L = 37, 57, 3
    x0 = (torch.randn(B, D, generator=g, dtype=torch.float64) * 0.5).requires_grad_()
    ws = [(torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5).requires_grad_() for _ in range(L)]
    bs = [(torch.randn(D, generator=g, dtype=torch.float64) * 0.1).requires_grad_() for _ in range(L)]
    gy = torch.randn(B, D, generator=g, dtype=torch.float64)
    y = V2.cross_v2_forward_torch(x0, ws, bs)
    y.backward(gy)
    y_np, dx0, gws, gbs = V2.cross_v2_numpy(x0.detach().numpy(), [w.detach().numpy() for w in ws],
                                            [b.detach().numpy() for b in bs], gy.numpy())
    np.testing.assert_allclose(y_np, y.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dx0, x0.grad.numpy(), rtol=1e-11, atol=1e-12)
    for l in range(L):
        np.testing.assert_allclose(gws[l], ws[l].grad.numpy(), rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(gbs[l], bs[l].grad.numpy(), rtol=1e-11, atol=1e-12)


def test_cross_v2_oracle_known_answer():
    """Hand-computed: D = 2, one layer, W = [[1, 2], [3, 4]], b = [0.5, -1], x0 = [1, 2]:
    u = W x0 + b = [5.5, 10], y = x0 * u + x0 = [6.5, 22]."""
    from oracle import cross_v2_oracle as V2
    x0 = torch.tensor([[1.0, 2.0]], dtype=torch.float64)
    y = V2.cross_v2_forward_torch(x0, [torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=torch.float64)],
                                  [torch.tensor([0.5, -1.0], dtype=torch.float64)])
    assert y.tolist() == [[6.5, 22.0]]
