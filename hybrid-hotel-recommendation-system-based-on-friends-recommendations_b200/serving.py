"""The ranking call of the service (main.py:319-325) on B200, with host buffers.

``rank_candidates`` is the one-request call the FastAPI endpoint makes; ``RankingEngine`` is the
batched form (many requests x candidates) that streams host tensors through the GPU in row chunks
with copies and compute overlapped on separate CUDA streams.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


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
    returns the logits on the host.  Row chunks are double-buffered: chunk i+1 is copied
    host->device on a copy stream while chunk i is scored; results go back device->host on the
    compute stream.  Pinned host tensors make the copies asynchronous."""

    def __init__(self, model, chunk_rows: int = 1 << 20, n_buffers: int = 2):
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
        for b in self.bufs:
            b["free"].record(compute)
        for i, r0 in enumerate(range(0, n, self.chunk_rows)):
            r1 = min(n, r0 + self.chunk_rows)
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
