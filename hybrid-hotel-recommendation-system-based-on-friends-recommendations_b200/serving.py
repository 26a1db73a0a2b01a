"""The ranking call of the service (main.py:319-325) on B200, with host buffers.

``rank_candidates`` is the one-request call the FastAPI endpoint makes; ``RankingEngine`` is the
batched form (many requests x candidates) that streams host tensors through the GPU in row chunks
with copies and compute overlapped on separate CUDA streams.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _cabi as C


def rank_candidates(model, X_user, X_item, X_cat, X_num) -> np.ndarray:
    """scores for one request's candidates -- the body of main.py:320-324.

    Inputs are the CPU tensors ``preprocess_for_ranking`` returns (main.py:221-230); they are moved
    to the model's device exactly like ``.to(ml_artifacts['device'])`` does at main.py:321-322.
    Returns ``preds.cpu().numpy()`` (0-d when there is a single candidate, as in the reference)."""
    dev = next(model.parameters()).device
    with torch.no_grad():
        preds = model(X_user.to(dev), X_item.to(dev), X_cat.to(dev), X_num.to(dev))
    return preds.cpu().numpy()


def sort_scored(scores: np.ndarray, item_ids):
    """sorted(zip(scores, item_ids), key=score, reverse=True) (main.py:325): stable, descending."""
    order = np.argsort(-np.asarray(scores, dtype=np.float64).reshape(-1), kind="stable")
    ids = np.asarray(list(item_ids))
    return [(float(np.asarray(scores).reshape(-1)[i]), ids[i].item()) for i in order]


class RankingEngine:
    """Streams [rows] of (user, item, cat, num) HOST tensors through ``model`` (eval mode) and
    returns the logits on the host.  Row chunks go through a ring of device buffers: chunks i+1, i+2 are
    copied host->device on a copy stream while chunk i is scored (a third buffer absorbs the jitter of either side: copy and
    compute take about the same time per chunk on a PCIe 5 host); results go back device->host on the
    compute stream.  Pinned host tensors make the copies asynchronous."""

    def __init__(self, model, chunk_rows: int = 1 << 20, n_buffers: int = 3):
        self.model = model
        self.dev = next(model.parameters()).device
        self.chunk_rows = int(chunk_rows)
        s = model._shape
        n_cat, n_num = len(s["cat_rows"]), s["n_num"]
        self.bufs = []
        for _ in range(n_buffers):
            self.bufs.append(dict(
                user=torch.empty(self.chunk_rows, dtype=torch.int64, device=self.dev),
                item=torch.empty(self.chunk_rows, dtype=torch.int64, device=self.dev),
                cat=torch.empty((self.chunk_rows, n_cat), dtype=torch.int64, device=self.dev),
                num=torch.empty((self.chunk_rows, n_num), dtype=torch.float32, device=self.dev),
                free=torch.cuda.Event(), ready=torch.cuda.Event()))
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.bytes_per_row_h2d = 8 + 8 + 8 * n_cat + 4 * n_num
        self.bytes_per_row_d2h = 4

    @torch.no_grad()
    def score(self, user_ids: torch.Tensor, item_ids: torch.Tensor, cat: torch.Tensor, num: torch.Tensor,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert not self.model.training, "RankingEngine scores in eval() mode (main.py:265)"
        n = user_ids.numel()
        if out is None:
            out = torch.empty(n, dtype=torch.float32, pin_memory=True)
        compute = torch.cuda.current_stream(self.dev)
        deferred, self.model.defer_eval_checks = self.model.defer_eval_checks, True     # one flag read-back per score()
        for b in self.bufs:
            b["free"].record(compute)
        # chunk schedule: the first chunks are short (1/8, 1/4, 1/2 of a chunk) so that scoring starts after a 0.2 ms copy
        # instead of a 1.5 ms one; from then on the copy of chunk i + 1 hides behind the scoring of chunk i
        bounds, r0, ramp = [], 0, 8
        while r0 < n:
            step = max(1, self.chunk_rows // ramp)
            bounds.append((r0, min(n, r0 + step)))
            r0 += step
            ramp = max(1, ramp // 2)
        for i, (r0, r1) in enumerate(bounds):
            rows = r1 - r0
            b = self.bufs[i % len(self.bufs)]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(b["free"])
                b["user"][:rows].copy_(user_ids[r0:r1], non_blocking=True)
                b["item"][:rows].copy_(item_ids[r0:r1], non_blocking=True)
                b["cat"][:rows].copy_(cat[r0:r1], non_blocking=True)
                b["num"][:rows].copy_(num[r0:r1], non_blocking=True)
                b["ready"].record(self.copy_stream)
            compute.wait_event(b["ready"])
            logits = self.model(b["user"][:rows], b["item"][:rows], b["cat"][:rows], b["num"][:rows])
            out[r0:r1].copy_(logits.reshape(-1), non_blocking=True)
            b["free"].record(compute)
        compute.synchronize()
        self.model.defer_eval_checks = deferred
        if self.model.check_eval_flags():
            # some chunk left the fp16 range of the fused tower: score everything again on the tf32x3 kernels
            prec, self.model.precision = self.model.precision, "tf32x3"
            try:
                return self.score(user_ids, item_ids, cat, num, out)
            finally:
                self.model.precision = prec
        return out


def rerank_with_mmr(ranked_items_with_scores, lambda_param: float, top_k: int = 20, *, item_embeddings: torch.Tensor,
                    item_id_mapping) -> list:
    """Drop-in for ``rerank_with_mmr`` (main.py:133-169) on the GPU: same arguments (the ranked ``[(score, item_id)]``
    list, ``lambda_param``, ``top_k``) plus the two objects the reference reads from ``ml_artifacts``
    (``item_embeddings`` as a CUDA tensor [n_items, d], ``artifacts['item_id_mapping']``).  Returns the re-ranked item ids."""
    return rerank_with_mmr_batch([ranked_items_with_scores], lambda_param, top_k, item_embeddings=item_embeddings,
                                 item_id_mapping=item_id_mapping)[0]


def rerank_with_mmr_batch(requests, lambda_param: float, top_k: int = 20, *, item_embeddings: torch.Tensor,
                          item_id_mapping) -> list:
    """Many requests in one launch (one CTA per request).  ``requests``: list of ranked ``[(score, item_id)]`` lists."""
    from . import _cabi as C
    C.require_cuda(item_embeddings)
    emb = item_embeddings.to(torch.float32).contiguous()
    dev = emb.device
    offsets, scores, idx = [0], [], []
    for req in requests:
        for score, item_id in req:
            scores.append(float(score))
            j = item_id_mapping.get(item_id)
            idx.append(-1 if j is None else int(j))
        offsets.append(len(scores))
    n_req = len(requests)
    if n_req == 0:
        return []
    max_c = max(b - a for a, b in zip(offsets[:-1], offsets[1:]))
    s_dev = torch.tensor(scores, dtype=torch.float32, device=dev)
    i_dev = torch.tensor(idx, dtype=torch.int64, device=dev)
    o_dev = torch.tensor(offsets, dtype=torch.int32, device=dev)
    order = torch.empty((n_req, top_k), dtype=torch.int32, device=dev)
    count = torch.empty(n_req, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().dcnr_mmr_rerank(C.ptr(emb), emb.shape[0], emb.shape[1], C.ptr(s_dev), C.ptr(i_dev), C.ptr(o_dev), n_req,
                                        float(lambda_param), int(top_k), int(max_c), C.ptr(order), C.ptr(count), C.stream()))
    order, count = order.cpu().numpy(), count.cpu().numpy()
    return [[requests[r][int(p)][1] for p in order[r, :count[r]]] for r in range(n_req)]


def expand_candidates(nn_model, item_embeddings: torch.Tensor, positive_rows, n_neighbors: int = 11):
    """The per-positive-hotel kNN loop of ``_generate_candidates`` (main.py:196-203) as ONE batched query: for every
    internal row in ``positive_rows`` the ``n_neighbors - 1`` nearest catalog rows with position 0 (the hotel itself)
    dropped, exactly what ``indices.squeeze()[1:]`` keeps at main.py:201.  Returns an int64 array [len(positive_rows),
    n_neighbors - 1] of internal rows; the caller maps them through ``reverse_item_map`` and unions them into the
    candidate set like the reference does."""
    rows = torch.as_tensor(positive_rows, dtype=torch.int64, device=item_embeddings.device).reshape(-1)
    if rows.numel() == 0:
        return np.empty((0, max(n_neighbors - 1, 0)), dtype=np.int64)
    _, ind = nn_model.kneighbors_tensor(item_embeddings[rows].to(torch.float32), n_neighbors)
    return ind[:, 1:].cpu().numpy()


# ------------------------------------------------------------------------------------------------
# Artifacts on disk (SURVEY 8b "Artifacts" row, 8f-4): the five files train.py:391-396 writes and
# main.py:256-269 reads, unchanged.
# ------------------------------------------------------------------------------------------------
ARTIFACT_FILES = ("final_dcn_model.pth", "artifacts.gz", "item_embeddings.npy", "best_params.gz", "model_dims.gz")


def save_ml_artifacts(artifacts_dir: str, model, artifacts: dict, best_params: dict, model_dims) -> None:
    """What train.py:391-396 does after training: state_dict, the preprocessing dict
    (user_id_mapping, item_id_mapping, scaler, cat_encoders, numerical_cols, categorical_cols; train.py:80-84),
    the item-embedding matrix, the hyper-parameters and (n_users, n_items, cat_dims, n_num_features)."""
    import os
    import joblib
    os.makedirs(artifacts_dir, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(artifacts_dir, "final_dcn_model.pth"))
    joblib.dump(artifacts, os.path.join(artifacts_dir, "artifacts.gz"))
    np.save(os.path.join(artifacts_dir, "item_embeddings.npy"), model.item_embedding.weight.detach().cpu().numpy())
    joblib.dump(best_params, os.path.join(artifacts_dir, "best_params.gz"))
    joblib.dump(model_dims, os.path.join(artifacts_dir, "model_dims.gz"))


def load_ml_artifacts(artifacts_dir: str = "artifacts", device="cuda", precision: Optional[str] = None) -> dict:
    """main.py:256-272 on the GPU: returns the reference's ``ml_artifacts`` entries that belong to this path --
    'device', 'artifacts', 'item_embeddings' (numpy, as the reference keeps it), 'final_model' (eval mode, on
    ``device``), 'nn_model' (fitted ``NearestNeighbors(n_neighbors=16, metric='cosine', algorithm='brute')``) and
    'reverse_item_map' -- plus 'item_embeddings_device' for the MMR / candidate-expansion calls.  The CSV frames
    ('main_df', 'friendships_df') stay with the service's pandas code (out of scope)."""
    import os
    import joblib
    from .knn import NearestNeighbors
    from .model import DCN_RecSys
    device = torch.device(device)
    out = {"device": device}
    out["artifacts"] = joblib.load(os.path.join(artifacts_dir, "artifacts.gz"))
    model_dims = joblib.load(os.path.join(artifacts_dir, "model_dims.gz"))
    best_params = joblib.load(os.path.join(artifacts_dir, "best_params.gz"))
    out["item_embeddings"] = np.load(os.path.join(artifacts_dir, "item_embeddings.npy"))
    n_users, n_items, cat_dims, n_num_features = model_dims
    model = DCN_RecSys(n_users, n_items, cat_dims, n_num_features, best_params)
    model.load_state_dict(torch.load(os.path.join(artifacts_dir, "final_dcn_model.pth"), map_location="cpu"))
    if precision is not None:
        model.precision = precision
    model.to(device)
    model.eval()
    out["final_model"] = model
    nn_model = NearestNeighbors(n_neighbors=16, metric="cosine", algorithm="brute")
    nn_model.fit(out["item_embeddings"])
    out["nn_model"] = nn_model
    out["reverse_item_map"] = {v: k for k, v in out["artifacts"]["item_id_mapping"].items()}
    out["item_embeddings_device"] = torch.from_numpy(out["item_embeddings"]).to(device)
    return out


# ------------------------------------------------------------------------------------------------
# Feature prep for the ranking call on the device (main.py:215-230)
# ------------------------------------------------------------------------------------------------
class RankingFeatures:
    """Per-hotel ranking features resident on the GPU.

    ``preprocess_for_ranking`` (main.py:215-230) rebuilds, per request and in pandas, the model inputs of the
    candidate hotels: internal user id (unknown -> len(map) // 2), internal item ids (unknown -> 0), categorical
    codes through ``cat_encoders`` (unknown -> 0) and the MinMax-scaled numerics.  Everything except the user id is a
    function of the hotel row, so it is computed ONCE here for the de-duplicated hotel frame (first row per
    ``item_id``, like ``drop_duplicates(subset=['item_id'])`` at main.py:314) with the same pandas / scikit-learn
    calls, and kept as device tensors; a request is then an index gather."""

    def __init__(self, artifacts: dict, hotels_df, device="cuda"):
        import pandas as pd  # noqa: F401  (the frame is pandas, like the reference's main_df)
        df = hotels_df.drop_duplicates(subset=["item_id"]).reset_index(drop=True)
        self.artifacts = artifacts
        self.device = torch.device(device)
        self.row_of_item = {int(h): r for r, h in enumerate(df["item_id"].tolist())}
        item_enc = df["item_id"].map(artifacts["item_id_mapping"]).fillna(0)
        self.item_internal = torch.tensor(item_enc.values, dtype=torch.long, device=self.device)
        cat = {}
        for col, encoder in artifacts["cat_encoders"].items():
            cat[col] = df[col].map(encoder).fillna(0).values
        cat_arr = np.stack([cat[c] for c in artifacts["cat_encoders"]], axis=1) if cat else np.zeros((len(df), 0))
        self.cat_codes = torch.tensor(cat_arr, dtype=torch.long, device=self.device)
        num_scaled = artifacts["scaler"].transform(df[artifacts["numerical_cols"]])
        self.num_scaled = torch.tensor(np.asarray(num_scaled), dtype=torch.float32, device=self.device)

    def internal_user_id(self, user_id: int) -> int:
        m = self.artifacts["user_id_mapping"]
        return m.get(user_id, len(m) // 2)

    def rows(self, hotel_ids) -> torch.Tensor:
        """Feature-table rows of raw hotel ids (the order of ``hotel_ids`` is kept)."""
        return torch.tensor([self.row_of_item[int(h)] for h in hotel_ids], dtype=torch.long, device=self.device)

    def preprocess_for_ranking(self, hotel_ids, user_id: int):
        """(X_collab_user, X_collab_item, X_cat, X_num) as CUDA tensors -- the tuple main.py:215-230 returns."""
        r = self.rows(hotel_ids)
        x_user = torch.full((r.numel(),), self.internal_user_id(user_id), dtype=torch.long, device=self.device)
        return x_user, self.item_internal[r], self.cat_codes[r], self.num_scaled[r]


class ItemFusedRanker:
    """The ranking forward with the feature prep folded into the embedding gather.

    The categorical embedding row and the numerics of a candidate depend only on its hotel row, so
    ``cat_embeddings[i].weight[code_i(hotel)]`` and ``num(hotel)`` are pre-gathered into per-hotel tables and the
    kernel's categorical segments are pointed at them: a request needs (user id, item id, hotel row) per candidate --
    24 bytes instead of 76 -- and ``x0`` (hence every logit) is bit-identical to the explicit-feature call."""

    def __init__(self, model, features: RankingFeatures):
        if model.training:
            raise RuntimeError("ItemFusedRanker serves an eval() model")
        self.model, self.features = model, features
        n_cat, n_num = features.cat_codes.shape[1], features.num_scaled.shape[1]
        if n_cat + (1 if n_num else 0) > 8:
            raise ValueError("too many categorical tables to fold the numerics in (DCNR_MAX_CAT = 8)")
        with torch.no_grad():
            self.tables = [model.cat_embeddings[i].weight.detach()[features.cat_codes[:, i]].contiguous() for i in range(n_cat)]
            if n_num:
                self.tables.append(features.num_scaled.contiguous())
        self.n_rows = features.cat_codes.shape[0]

    def score(self, user_internal: int, item_internal: torch.Tensor, hotel_rows: torch.Tensor) -> torch.Tensor:
        m = self.model
        dev = item_internal.device
        B = item_internal.numel()
        dims, ps = m._dims(), m._param_struct()
        dims.n_cat, dims.n_num = len(self.tables), 0
        for i, t in enumerate(self.tables):
            dims.cat_rows[i], dims.cat_width[i] = self.n_rows, t.shape[1]
            ps.cat_table[i] = C.ptr(t)
        users = torch.full((B,), int(user_internal), dtype=torch.long, device=dev)
        cat = hotel_rows.reshape(B, 1).expand(B, len(self.tables)).contiguous()
        batch = C.Batch(C.ptr(users), C.ptr(item_internal.contiguous()), C.ptr(cat), None, B)
        ws = torch.empty(C.lib().dcnr_workspace_bytes(dims, B, 0), dtype=torch.uint8, device=dev)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        C.check(C.lib().dcnr_forward_eval(dims, ps, batch, C.ptr(logits), C.ptr(ws), ws.numel(), C.stream()))
        return logits

    def rank(self, hotel_ids, user_id: int):
        """main.py:319-325 for one request: scores of the candidate hotels, sorted descending with the reference's
        stable python sort (``sort_scored``)."""
        f = self.features
        r = f.rows(hotel_ids)
        scores = self.score(f.internal_user_id(user_id), f.item_internal[r], r).cpu().numpy()
        return sort_scored(scores, list(hotel_ids))
