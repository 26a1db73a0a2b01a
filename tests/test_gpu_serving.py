"""GPU parity of the rows after the ranking forward through the C ABI:
MMR re-rank vs the reference's own outputs (golden) and vs the oracle on larger random requests;
batched candidate expansion vs the reference's one-query-per-positive-hotel loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
GOLDEN = GOLD


@pytest.mark.parametrize("name", ["c300", "c40_unmapped", "c12"])
def test_mmr_matches_reference_golden(name):
    import dcnr_b200
    z = np.load(os.path.join(GOLD, f"mmr_{name}.npz"))
    E = torch.from_numpy(z["E"]).cuda()
    C = len(z["scores"])
    ids = [int(i) * 7 + 3 if i >= 0 else -1000 - c for c, i in enumerate(z["emb_idx"])]
    mapping = {int(i) * 7 + 3: int(i) for i in range(E.shape[0])}
    ranked = [(float(z["scores"][c]), ids[c]) for c in range(C)]
    got = dcnr_b200.serving.rerank_with_mmr(ranked, float(z["lam"]), int(z["top_k"]), item_embeddings=E, item_id_mapping=mapping)
    assert got == [ids[p] for p in z["order"]]


def test_mmr_batch_matches_oracle():
    import dcnr_b200
    from oracle import mmr_oracle
    rng = np.random.default_rng(3)
    E = rng.standard_normal((20000, 16)).astype(np.float32)
    E[77] = 0.0                                             # zero vector: norm 0 -> 1 like sklearn's normalize
    mapping = {i: i for i in range(E.shape[0])}
    reqs, want = [], []
    for r in range(64):
        C = int(rng.integers(1, 700))
        idx = rng.choice(E.shape[0], size=C, replace=False)
        if r % 5 == 0:
            idx[0] = 77
        sc = np.sort(rng.standard_normal(C).astype(np.float32))[::-1].copy()
        if r % 7 == 0 and C > 4:
            sc[2] = sc[1]                                   # exact score tie: the earlier candidate must win
        reqs.append([(sc[c], int(idx[c])) for c in range(C)])
        want.append([int(idx[p]) for p in mmr_oracle.mmr_rerank(E, sc, idx.astype(np.int64), 0.7, 20)])
    got = dcnr_b200.serving.rerank_with_mmr_batch(reqs, 0.7, 20, item_embeddings=torch.from_numpy(E).cuda(), item_id_mapping=mapping)
    assert got == want


def test_expand_candidates_equals_per_hotel_queries():
    """main.py:196-203: one kneighbors(vec, 11) per positive hotel, position 0 dropped -- batched here."""
    import dcnr_b200
    from oracle import knn_oracle
    rng = np.random.default_rng(11)
    E = rng.standard_normal((30000, 16)).astype(np.float32)
    positives = rng.choice(E.shape[0], size=37, replace=False)
    nn_model = dcnr_b200.NearestNeighbors(n_neighbors=16, metric="cosine", algorithm="brute").fit(E)
    got = dcnr_b200.serving.expand_candidates(nn_model, torch.from_numpy(E).cuda(), positives, 11)
    ref = knn_oracle.OracleNearestNeighbors().fit(E)
    for r, row in enumerate(positives):
        _, ind = ref.kneighbors(E[row].reshape(1, -1), n_neighbors=11)
        assert np.array_equal(got[r], ind[0][1:])
    assert set(got.reshape(-1)) == set(np.concatenate([ref.kneighbors(E[p].reshape(1, -1), n_neighbors=11)[1][0][1:] for p in positives]))


# ------------------------------------------------------------------------------------------------
# artifacts on disk + device feature prep (SURVEY 8b artifacts row, 8f-4)
# ------------------------------------------------------------------------------------------------
def _serving_fixture(tmp_path):
    import dcnr_b200
    import joblib
    import pandas as pd
    blob = joblib.load(os.path.join(GOLDEN, "preprocess_ranking_inputs.gz"))
    df, artifacts = blob["frame"], blob["artifacts"]
    cat_dims = {c: len(e) for c, e in artifacts["cat_encoders"].items()}
    dims = (len(artifacts["user_id_mapping"]), len(artifacts["item_id_mapping"]), cat_dims, len(artifacts["numerical_cols"]))
    best = dict(emb_dim=16, hidden_dim=64, n_cross_layers=2, dropout=0.3, n_res_blocks=2)
    torch.manual_seed(5)
    model = dcnr_b200.DCN_RecSys(*dims, best)
    with torch.no_grad():                       # non-trivial running statistics
        for blk in model.res_blocks:
            for bn in (blk.bn1, blk.bn2):
                bn.running_mean.normal_(0, 0.3)
                bn.running_var.uniform_(0.5, 1.5)
    dcnr_b200.serving.save_ml_artifacts(str(tmp_path), model, artifacts, best, dims)
    return df, artifacts, model


@pytest.mark.gpu
def test_load_ml_artifacts_round_trip_and_fused_item_tables(tmp_path):
    import dcnr_b200
    df, artifacts, model_cpu = _serving_fixture(tmp_path)
    ml = dcnr_b200.serving.load_ml_artifacts(str(tmp_path), device="cuda", precision="fp32")
    model = ml["final_model"]
    assert not model.training and next(model.parameters()).is_cuda
    for k, v in model_cpu.state_dict().items():
        assert torch.equal(model.state_dict()[k].cpu(), v), k
    assert ml["reverse_item_map"] == {v: k for k, v in artifacts["item_id_mapping"].items()}
    # the kNN object answers like main.py:300-302 expects (position 0 is the query item itself)
    _, ind = ml["nn_model"].kneighbors(ml["item_embeddings"][3].reshape(1, -1), n_neighbors=6)
    assert ind.dtype == np.int64 and ind.shape == (1, 6) and ind[0, 0] == 3

    feats = dcnr_b200.serving.RankingFeatures(artifacts, df, device="cuda")
    hotels = list(dict.fromkeys(df["item_id"].tolist()))
    user_id = int(df["user_id"].iloc[0])
    xu, xi, xc, xn = feats.preprocess_for_ranking(hotels, user_id)
    with torch.no_grad():
        ref_scores = model(xu, xi, xc, xn)
    fused = dcnr_b200.serving.ItemFusedRanker(model, feats)
    r = feats.rows(hotels)
    got = fused.score(feats.internal_user_id(user_id), feats.item_internal[r], r)
    assert torch.equal(got, ref_scores)                               # same x0 rows -> bit-identical logits
    ranked = fused.rank(hotels, user_id)
    assert [h for _, h in ranked] == [h for _, h in dcnr_b200.serving.sort_scored(ref_scores.cpu().numpy(), hotels)]
    assert all(ranked[i][0] >= ranked[i + 1][0] for i in range(len(ranked) - 1))


def test_graphed_train_step_equals_eager_loop_with_stock_optimizer():
    """training.GraphedTrainStep: N steps of the reference's loop (zero_grad -> forward -> BCE -> backward ->
    optimizer.step, train.py:218-226) replayed from one CUDA graph leave the same parameters, BatchNorm buffers and
    num_batches_tracked as the eager loop -- capture() itself must not move the running statistics, and zero_grad()'s
    default set_to_none=True must not disconnect the optimizer from the graph's gradient buffers."""
    import dcnr_b200
    from oracle import dcnr_oracle as orc
    from tests.helpers import synth_inputs
    n_users, n_items, cat_dims, n_num = 3000, 900, {"city": 100, "hotel_type": 6}, 11
    params = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0)
    state = orc.make_state(n_users, n_items, cat_dims, n_num, params, seed=4, emb_scale=0.1, randomize_bn=True)
    B, steps = 2048, 4
    batches = [tuple(t.cuda() for t in synth_inputs(n_users, n_items, cat_dims, n_num, B, seed=50 + s)) for s in range(steps)]

    def fresh():
        m = dcnr_b200.DCN_RecSys(n_users, n_items, cat_dims, n_num, params, precision="tf32x3")
        m.load_state_dict(state)
        return m.cuda().train()

    # the eager loop as training.py documents it (forward, fused BCE, backward with the loss gradient, optimizer step); the graph
    # replays exactly these kernels, so the two runs must agree BIT FOR BIT
    eager = fresh()
    opt = torch.optim.SGD(eager.parameters(), lr=0.05, momentum=0.9)
    for u, i, c, x, y in batches:
        opt.zero_grad()
        logits = eager(u, i, c, x)
        _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
        logits.backward(gradient=dl)
        opt.step()

    graphed = fresh()
    opt_g = torch.optim.SGD(graphed.parameters(), lr=0.05, momentum=0.9)
    gs = dcnr_b200.training.GraphedTrainStep(graphed, B)
    gs.load(*batches[0])
    gs.capture()
    for n, b in graphed.named_buffers():                      # capture left the model as it found it
        assert torch.equal(b.cpu(), state[n]), n
    for u, i, c, x, y in batches:
        opt_g.zero_grad()                                     # set_to_none=True: detaches .grad
        gs(u, i, c, x, y)
        opt_g.step()
    for (n, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert not torch.equal(pe.detach().cpu(), state[n]) or ".bias" in n or "cross" in n, n      # the optimizer really stepped
        assert torch.equal(pe, pg), n
    for (n, be), (_, bg) in zip(eager.named_buffers(), graphed.named_buffers()):
        assert torch.equal(be, bg), n
        if not be.dtype.is_floating_point:
            assert int(be) == steps, n


def test_row_sharded_model_on_one_rank_equals_the_plain_model():
    """RowShardedDCN with a one-rank communicator (no exchange: the fused two-table lookup gathers locally) gives the logits
    and gradients of DCN_RecSys with the full tables."""
    import dcnr_b200
    from dcnr_b200 import distributed as D
    from oracle import dcnr_oracle as orc
    from tests.helpers import synth_inputs
    nu, ni, cat, nn_ = 3000, 700, {"city": 100, "hotel_type": 6}, 11
    params = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0)
    state = orc.make_state(nu, ni, cat, nn_, params, seed=4, emb_scale=0.1, randomize_bn=True)
    u, i, c, x, _ = synth_inputs(nu, ni, cat, nn_, 1500, seed=8, zipf=True)
    gl = torch.randn(1500, generator=torch.Generator().manual_seed(1)) / 1500
    dev = torch.device("cuda")
    full = dcnr_b200.DCN_RecSys(nu, ni, cat, nn_, params, precision="fp32")
    full.load_state_dict(state); full.to(dev).train()
    lo_full = full(u.to(dev), i.to(dev), c.to(dev), x.to(dev))
    lo_full.backward(gradient=gl.to(dev))
    comm = D.Communicator()
    assert comm.world == 1
    sh = D.RowShardedDCN(nu, ni, cat, nn_, params, comm, precision="fp32", device=dev)
    core_state = {k: v for k, v in state.items() if not k.startswith(("user_embedding", "item_embedding"))}
    sh.core.load_state_dict({**core_state, "user_embedding.weight": torch.zeros(1, 16), "item_embedding.weight": torch.zeros(1, 16)})
    sh.core.to(dev)
    sh.user_table.load_full(state["user_embedding.weight"]); sh.item_table.load_full(state["item_embedding.weight"])
    sh.train()
    lo = sh(u.to(dev), i.to(dev), c.to(dev), x.to(dev))
    lo.backward(gradient=gl.to(dev))

    def err(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    assert err(lo.detach(), lo_full.detach()) < 1e-6
    assert err(sh.user_table.weight.grad, full.user_embedding.weight.grad) < 2e-6
    assert err(sh.item_table.weight.grad, full.item_embedding.weight.grad) < 2e-6
    assert err(sh.core.initial_deep_layer.weight.grad, full.initial_deep_layer.weight.grad) < 1e-5
    comm.close()
