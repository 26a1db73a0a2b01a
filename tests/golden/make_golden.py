"""Generates tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run once here (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It imports /root/reference/main.py (whose model classes are the service's copy of
train.py:90-170) and scikit-learn's NearestNeighbors exactly as main.py:268-269 builds it,
feeds them small seeded inputs and stores inputs + state_dict + outputs + gradients.
Nothing from the reference's source is copied into the repo -- only numeric vectors.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("REF_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import main as ref_main  # noqa: E402  (reference module; import has no side effects beyond logging/seed)
from sklearn.neighbors import NearestNeighbors  # noqa: E402


def _pack_state(sd):
    return {"sd::" + k: v.detach().cpu().numpy() for k, v in sd.items()}


def model_case(name, n_users, n_items, cat_dims, n_num, params, B, seed, emb_scale=1.0, randomize_bn=False):
    torch.manual_seed(seed)
    model = ref_main.DCN_RecSys(n_users, n_items, cat_dims, n_num, params)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        if emb_scale != 1.0:
            model.user_embedding.weight.mul_(emb_scale)
            model.item_embedding.weight.mul_(emb_scale)
            for e in model.cat_embeddings:
                e.weight.mul_(emb_scale)
        if randomize_bn:
            for blk in model.res_blocks:
                for bn in (blk.bn1, blk.bn2):
                    bn.weight.copy_(0.5 + torch.rand(bn.weight.shape, generator=g))
                    bn.bias.copy_(torch.randn(bn.bias.shape, generator=g) * 0.2)
                    bn.running_mean.copy_(torch.randn(bn.bias.shape, generator=g) * 0.3)
                    bn.running_var.copy_(0.5 + torch.rand(bn.bias.shape, generator=g))
            for cl in model.cross_network:
                cl.b.copy_(torch.randn(cl.b.shape, generator=g) * 0.05)
    user_ids = torch.randint(0, n_users, (B,), generator=g)
    item_ids = torch.randint(0, n_items, (B,), generator=g)
    # duplicate-heavy ids in the second half so the scatter sees long segments
    user_ids[B // 2:] = user_ids[B // 2:] % max(1, n_users // 8)
    cat = torch.stack([torch.randint(0, n, (B,), generator=g) for n in cat_dims.values()], dim=1)
    num = torch.rand(B, n_num, generator=g)
    labels = (torch.rand(B, generator=g) < 0.3).float()
    grad_logits = torch.randn(B, generator=g) / B

    sd0 = {k: v.clone() for k, v in model.state_dict().items()}

    # eval-mode forward (the ranking call, main.py:320-322)
    model.eval()
    with torch.no_grad():
        logits_eval = model(user_ids, item_ids, cat, num)

    # train-mode forward + backward with an explicit upstream gradient (train.py:223-225)
    model.train()
    model.zero_grad()
    logits_train = model(user_ids, item_ids, cat, num)
    logits_train.backward(gradient=grad_logits)
    grads = {"grad::" + k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
    sd1 = {k: v.clone() for k, v in model.state_dict().items()}   # running stats after one train step

    # same through the loss (train.py:224-225)
    model.load_state_dict(sd0)
    model.zero_grad()
    loss = torch.nn.BCEWithLogitsLoss()(model(user_ids, item_ids, cat, num), labels)
    loss.backward()
    grads_loss = {"lossgrad::" + k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}

    # float64 arbiter
    model.load_state_dict(sd0)
    m64 = ref_main.DCN_RecSys(n_users, n_items, cat_dims, n_num, params).double()
    m64.load_state_dict(sd0)
    m64.train()
    logits64 = m64(user_ids, item_ids, cat, num.double())

    out = dict(
        meta=np.array([n_users, n_items, n_num, B, params["emb_dim"], params["hidden_dim"],
                       params["n_cross_layers"], params.get("n_res_blocks", 2)], dtype=np.int64),
        cat_dims=np.array(list(cat_dims.values()), dtype=np.int64),
        user_ids=user_ids.numpy(), item_ids=item_ids.numpy(), cat=cat.numpy(), num=num.numpy(),
        labels=labels.numpy(), grad_logits=grad_logits.numpy(),
        logits_eval=logits_eval.numpy(), logits_train=logits_train.detach().numpy(),
        logits_train_f64=logits64.detach().numpy(), loss=np.array(loss.item(), dtype=np.float64),
    )
    out.update(_pack_state(sd0))
    out.update({"after::" + k: v.numpy() for k, v in sd1.items() if "running" in k or "num_batches" in k})
    out.update(grads)
    if n_users * params['emb_dim'] + params['hidden_dim'] ** 2 < 20000:   # keep the big fixtures small
        out.update(grads_loss)
    path = os.path.join(HERE, f"model_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; loss", loss.item())


def cross_case():
    torch.manual_seed(7)
    layer = ref_main.CrossLayer(57)
    with torch.no_grad():
        layer.b.copy_(torch.randn(57) * 0.1)
    x = torch.randn(33, 57, requires_grad=True)
    y = layer(x)
    g = torch.randn(33, 57)
    y.backward(g)
    np.savez_compressed(os.path.join(HERE, "cross_layer.npz"), x=x.detach().numpy(), w=layer.w.weight.detach().numpy(),
                        b=layer.b.detach().numpy(), y=y.detach().numpy(), g=g.numpy(), gx=x.grad.numpy(),
                        gw=layer.w.weight.grad.numpy(), gb=layer.b.grad.numpy())


def resblock_case():
    torch.manual_seed(11)
    blk = ref_main.ResBlock(64, 0.0)
    with torch.no_grad():
        for bn in (blk.bn1, blk.bn2):
            bn.weight.copy_(0.5 + torch.rand(64)); bn.bias.copy_(torch.randn(64) * 0.2)
    x = torch.randn(48, 64, requires_grad=True)
    sd0 = {k: v.clone() for k, v in blk.state_dict().items()}      # BEFORE the train step updates running stats
    blk.train()
    y = blk(x)
    g = torch.randn(48, 64)
    y.backward(g)
    d = dict(x=x.detach().numpy(), y=y.detach().numpy(), g=g.numpy(), gx=x.grad.numpy())
    d.update({"sd::" + k: v.numpy() for k, v in sd0.items()})
    d.update({"grad::" + k: p.grad.numpy() for k, p in blk.named_parameters()})
    blk.eval()
    with torch.no_grad():
        d["y_eval"] = blk(x.detach()).numpy()
    np.savez_compressed(os.path.join(HERE, "res_block.npz"), **d)


def knn_case(name, n, d, nq, k, seed):
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((n, d)).astype(np.float32)
    q_rows = rng.integers(0, n, size=nq)
    Q = E[q_rows].copy()                                   # the reference queries with catalog rows (main.py:199,299)
    nn_model = NearestNeighbors(n_neighbors=16, metric="cosine", algorithm="brute")   # main.py:268
    nn_model.fit(E)                                                                      # main.py:269
    dists, inds = [], []
    for i in range(nq):                                    # one [1,d] query per call, as main.py:200,300
        dd, ii = nn_model.kneighbors(Q[i].reshape(1, -1), n_neighbors=k)
        dists.append(dd[0]); inds.append(ii[0])
    np.savez_compressed(os.path.join(HERE, f"knn_{name}.npz"), E=E, Q=Q, q_rows=q_rows, k=np.array(k),
                        dist=np.stack(dists), ind=np.stack(inds))
    print("knn", name, "dist dtype", dists[0].dtype, "ind dtype", inds[0].dtype)


def mmr_case(name, n_items, d, C, lam, top_k, seed, unmapped=()):
    """rerank_with_mmr (main.py:133-169) on random embeddings / scores; external ids are idx * 7 + 3 so that the id
    mapping is exercised; `unmapped` candidate positions get ids the mapping does not know (skipped at main.py:150)."""
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((n_items, d)).astype(np.float32)
    idx = rng.choice(n_items, size=C, replace=False)
    scores = np.sort(rng.standard_normal(C).astype(np.float32))[::-1].copy()       # ranked order (main.py:325)
    ext = [int(i) * 7 + 3 for i in idx]
    for pos in unmapped:
        ext[pos] = -1000 - pos
    ref_main.ml_artifacts["item_embeddings"] = E
    ref_main.ml_artifacts["artifacts"] = {"item_id_mapping": {int(i) * 7 + 3: int(i) for i in range(n_items)}}
    ranked = [(scores[c], ext[c]) for c in range(C)]         # numpy float32 scores, as zip(scores, ids) yields at main.py:325
    chosen = ref_main.rerank_with_mmr(ranked, lam, top_k)
    pos_of = {e: c for c, e in enumerate(ext)}
    emb_idx = np.array([-1 if c in unmapped else int(idx[c]) for c in range(C)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"mmr_{name}.npz"), E=E, scores=scores, emb_idx=emb_idx, lam=np.array(lam),
                        top_k=np.array(top_k), order=np.array([pos_of[e] for e in chosen], dtype=np.int32))
    print("mmr", name, "selected", len(chosen))


def synth_hotels(n_hotels=40, seed=0):
    """A tiny frame in the hackathon_augmented_data.csv schema + the preprocessing artifacts
    train.py:40-84 would build from it (same calls: category codes, MinMaxScaler)."""
    import pandas as pd
    from sklearn.preprocessing import MinMaxScaler
    rng = np.random.default_rng(seed)
    cities = ["Kazan", "Moscow", "Sochi", "Ufa"]
    types = ["hotel", "hostel", "apart"]
    hotel_id = rng.permutation(np.arange(1000, 1000 + n_hotels))
    rows = []
    for h in hotel_id:
        hr = np.random.default_rng(int(h))
        base = dict(hotel_id=int(h), city=cities[int(h) % 4], hotel_type=types[int(h) % 3],
                    price_rub=float(np.exp(hr.normal(8, 0.5))), stars=int(hr.integers(0, 6)),
                    user_reviews_count=int(hr.integers(0, 500)),
                    rating_location=float(hr.uniform(1, 10)), rating_cleanliness=float(hr.uniform(1, 10)),
                    rating_food=float(hr.uniform(1, 10)), rating_service=float(hr.uniform(1, 10)))
        for _ in range(int(hr.integers(1, 4))):           # several reviews per hotel, hotel attributes repeated
            rows.append(dict(base, guest_id=int(rng.integers(1, 30)), rating_overall=float(rng.uniform(1, 10)),
                             was_booked=int(rng.integers(0, 2))))
    df = pd.DataFrame(rows)
    df.rename(columns={"guest_id": "user_id", "hotel_id": "item_id"}, inplace=True)
    # derived features exactly as main.py:244-250 / train.py:284-287
    df["price_per_star"] = (df["price_rub"] / df["stars"]).replace([np.inf, -np.inf], 0).fillna(0)
    df["cleanliness_vs_service"] = (df["rating_cleanliness"] / df["rating_service"]).replace([np.inf, -np.inf], 0).fillna(0)
    df["location_premium"] = df["rating_overall"] - df["rating_location"]
    categorical_cols = ["city", "hotel_type"]
    numerical_cols = ["price_rub", "stars", "user_reviews_count", "rating_overall", "rating_location", "rating_cleanliness",
                      "rating_food", "rating_service", "price_per_star", "cleanliness_vs_service", "location_premium"]
    # the training frame misses the last 5 hotels, one city and a few users: the serving path must map those to 0 / mid id
    train = df[~df["item_id"].isin(hotel_id[-5:]) & (df["city"] != "Ufa")]
    user_map = {o: i for i, o in enumerate(train["user_id"].unique())}
    item_map = {o: i for i, o in enumerate(train["item_id"].unique())}
    cat_encoders = {}
    for col in categorical_cols:
        cats = train[col].astype("category").cat.categories
        cat_encoders[col] = {c: i for i, c in enumerate(cats)}
    scaler = MinMaxScaler().fit(train[numerical_cols])
    artifacts = {"user_id_mapping": user_map, "item_id_mapping": item_map, "scaler": scaler, "cat_encoders": cat_encoders,
                 "numerical_cols": numerical_cols, "categorical_cols": categorical_cols}
    return df, artifacts


def preprocess_case():
    """Runs the reference's preprocess_for_ranking (main.py:215-230) on a de-duplicated candidate frame, as the
    endpoint does (main.py:313-319), for a known and an unknown user."""
    import joblib
    df, artifacts = synth_hotels()
    ref_main.ml_artifacts["artifacts"] = artifacts
    out = {}
    for tag, user_id, pick in (("known", int(df["user_id"].iloc[0]), slice(0, None, 2)), ("unknown", 987654, slice(1, None, 3))):
        cand = list(dict.fromkeys(df["item_id"].tolist()))[pick]
        items = df[df["item_id"].isin(cand)].drop_duplicates(subset=["item_id"])
        xu, xi, xc, xn = ref_main.preprocess_for_ranking(items, user_id)
        out[f"{tag}::user_id"] = np.int64(user_id)
        out[f"{tag}::hotel_ids"] = items["item_id"].values.astype(np.int64)
        out[f"{tag}::x_user"], out[f"{tag}::x_item"] = xu.numpy(), xi.numpy()
        out[f"{tag}::x_cat"], out[f"{tag}::x_num"] = xc.numpy(), xn.numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess_ranking.npz"), **out)
    # the frame is stored with joblib, not CSV: a text round trip moves some float64 values by an ulp
    joblib.dump({"frame": df, "artifacts": artifacts}, os.path.join(HERE, "preprocess_ranking_inputs.gz"))
    print("preprocess_ranking", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    P0 = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0,
              lr=1e-3, batch_size=512)   # extra keys must be ignored (train.py:186-192)
    model_case("p0", 300, 120, {"city": 100, "hotel_type": 6}, 11, P0, B=96, seed=42)
    model_case("p0_trained", 300, 120, {"city": 100, "hotel_type": 6}, 11, P0, B=96, seed=43,
               emb_scale=0.1, randomize_bn=True)
    PODD = dict(emb_dim=24, hidden_dim=96, n_cross_layers=5, n_res_blocks=3, dropout=0.0)
    model_case("odd", 50, 31, {"city": 17, "hotel_type": 3, "extra": 40}, 7, PODD, B=37, seed=44,
               emb_scale=0.3, randomize_bn=True)
    PMIN = dict(emb_dim=16, hidden_dim=32, n_cross_layers=1, dropout=0.0)  # n_res_blocks default 2 (train.py:134)
    model_case("min", 10, 9, {"city": 4, "hotel_type": 2}, 11, PMIN, B=5, seed=45, emb_scale=0.5)
    cross_case()
    resblock_case()
    knn_case("small", 3000, 16, 8, 11, 5)
    knn_case("k51", 5000, 16, 4, 51, 6)
    knn_case("d64", 2000, 64, 4, 11, 7)
    mmr_case("c300", 2000, 16, 300, 0.7, 20, 8)
    mmr_case("c40_unmapped", 500, 16, 40, 0.3, 20, 9, unmapped=(0, 5, 17))
    mmr_case("c12", 200, 64, 12, 0.5, 20, 10)
    preprocess_case()
