"""Drop-in for ``sklearn.neighbors.NearestNeighbors(metric='cosine', algorithm='brute')`` as the
reference uses it (main.py:268-269 fit; main.py:200 and :300 kneighbors), on B200.

``fit`` uploads the item-embedding catalog once and pre-normalises it (sklearn re-normalises the
whole catalog on every query); ``kneighbors`` runs the exact-fp32 scan + top-k kernels and returns
``(dist float32 [q,k], ind int64 [q,k])`` sorted by (distance ascending, index ascending) -- the
fully specified order of oracle/knn_oracle.c, which coincides with sklearn's on tie-free data.
A catalog can be a shard of a larger one (``index_base``); ``merge_shards`` combines per-shard
results so the answer is independent of the shard count.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi as C


class NearestNeighbors:
    def __init__(self, n_neighbors: int = 5, metric: str = "cosine", algorithm: str = "brute", device=None,
                 index_base: int = 0, **_ignored):
        if metric != "cosine":
            raise ValueError("dcnr_b200.NearestNeighbors implements metric='cosine' only (main.py:268)")
        if algorithm not in ("brute", "auto"):
            raise ValueError("only algorithm='brute' is on the reference's path")
        self.n_neighbors = n_neighbors
        self.metric, self.algorithm = metric, algorithm
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None
        self.index_base = int(index_base)
        self.tc_min_queries = 8             # query batches from this size on take the tensor-core shortlist path
        self._catalog_hat: Optional[torch.Tensor] = None

    # ---- fit ---------------------------------------------------------------------------------
    def fit(self, X, y=None):
        if self.device is None:
            raise RuntimeError("dcnr_b200.NearestNeighbors needs a CUDA device (no CPU fallback)")
        x = torch.as_tensor(np.asarray(X) if not torch.is_tensor(X) else X)
        if x.dim() != 2:
            raise ValueError(f"Expected 2D array, got {x.dim()}D array instead")
        x = x.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            C.check(C.lib().dcnr_knn_normalize(C.ptr(x), C.ptr(out), x.shape[0], x.shape[1], C.stream()))
        self._catalog_hat = out
        self.n_samples_fit_, self.n_features_in_ = x.shape[0], x.shape[1]
        return self

    # ---- query -------------------------------------------------------------------------------
    def kneighbors_tensor(self, Q: torch.Tensor, n_neighbors: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident variant: Q [q,d] float32 CUDA -> (dist [q,k] f32, ind [q,k] i64) CUDA tensors."""
        if self._catalog_hat is None:
            raise RuntimeError("This NearestNeighbors instance is not fitted yet.")
        k = self.n_neighbors if n_neighbors is None else int(n_neighbors)
        n, d = self._catalog_hat.shape
        if k > n and self.index_base == 0 and not getattr(self, "_allow_short", False):
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {n}")
        if k < 1:
            raise ValueError(f"Expected n_neighbors > 0. Got {k}")
        C.require_cuda(Q)
        Q = Q.reshape(-1, d).to(torch.float32).contiguous()
        nq = Q.shape[0]
        dist = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
        ind = torch.empty((nq, k), dtype=torch.int64, device=Q.device)
        with torch.cuda.device(Q.device):
            qhat = torch.empty_like(Q)
            C.check(C.lib().dcnr_knn_normalize(C.ptr(Q), C.ptr(qhat), nq, d, C.stream()))
            # Query batches: tensor-core shortlist + exact re-score (same results bit for bit; csrc/topk_tc.cu).  The status
            # word is read back once per call: a shortlist that overflowed (pathological duplicates) sends the batch down
            # the exact streaming path below.
            if (nq >= self.tc_min_queries and C.lib().dcnr_knn_tc_supported(n, d, nq, k)
                    and not torch.cuda.is_current_stream_capturing()):      # the status read-back is a sync
                ws = torch.empty(C.lib().dcnr_knn_tc_scratch_bytes(n, d, nq, k), dtype=torch.uint8, device=Q.device)
                status = torch.zeros(1, dtype=torch.int32, device=Q.device)
                C.check(C.lib().dcnr_knn_topk_tc(C.ptr(self._catalog_hat), n, d, C.ptr(qhat), nq, k, self.index_base,
                                                 C.ptr(dist), C.ptr(ind), C.ptr(ws), ws.numel(), C.ptr(status), C.stream()))
                if int(status.item()) == 0:
                    return dist, ind
            # k up to 256 per launch; query tiles of 1024 bound the scratch
            for q0 in range(0, nq, 1024):
                q1 = min(nq, q0 + 1024)
                nbytes = C.lib().dcnr_knn_scratch_bytes(n, d, q1 - q0, k)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=Q.device)
                C.check(C.lib().dcnr_knn_topk(C.ptr(self._catalog_hat), n, d, C.ptr(qhat[q0:q1]), q1 - q0, k,
                                              self.index_base, C.ptr(dist[q0:q1]), C.ptr(ind[q0:q1]), C.ptr(ws),
                                              ws.numel(), C.stream()))
        return dist, ind

    def kneighbors(self, X=None, n_neighbors: Optional[int] = None, return_distance: bool = True):
        """sklearn protocol: array-like [q,d] (or [d]) in, numpy (dist f32, ind i64) out (main.py:200,300)."""
        if X is None:
            raise ValueError("kneighbors(X=None) (query = training set) is not on the reference's path")
        if torch.is_tensor(X) and X.is_cuda:
            dist, ind = self.kneighbors_tensor(X, n_neighbors)
            return (dist, ind) if return_distance else ind
        q = np.asarray(X, dtype=np.float32)
        if q.ndim != 2:
            raise ValueError(f"Expected 2D array, got {q.ndim}D array instead")
        dist, ind = self.kneighbors_tensor(torch.from_numpy(np.ascontiguousarray(q)).to(self.device), n_neighbors)
        dist, ind = dist.cpu().numpy(), ind.cpu().numpy()
        return (dist, ind) if return_distance else ind


def merge_shards(dist_parts: torch.Tensor, ind_parts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_parts, q, k] per-shard results (global indices, (inf,-1) padded) -> [q, k] in the contract
    order.  The cross-GPU merge of the sharded-catalog configuration."""
    C.require_cuda(dist_parts, ind_parts)
    dist_parts = dist_parts.to(torch.float32).contiguous()
    ind_parts = ind_parts.to(torch.int64).contiguous()
    n_parts, nq, k = dist_parts.shape
    dist = torch.empty((nq, k), dtype=torch.float32, device=dist_parts.device)
    ind = torch.empty((nq, k), dtype=torch.int64, device=dist_parts.device)
    with torch.cuda.device(dist_parts.device):
        C.check(C.lib().dcnr_knn_merge(C.ptr(dist_parts), C.ptr(ind_parts), n_parts, nq, k, C.ptr(dist), C.ptr(ind),
                                       C.stream()))
    return dist, ind
