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
