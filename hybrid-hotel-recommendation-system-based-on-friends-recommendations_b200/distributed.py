"""Multi-GPU plumbing for the DCN-R path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

What shards and how:
  * ranking inference  -- requests are independent: contiguous request ranges per rank, weights
    replicated, NO collective on the data path (results are gathered by the caller);
  * training           -- data parallel: the batch is split, parameters replicated, gradients
    averaged with one flat-bucket all-reduce (dense parameters) plus one all-reduce per embedding
    table; BatchNorm statistics stay per-rank (the torch DDP default) unless the caller opts into
    ``sync_bn`` (not implemented yet -- see DESIGN.md);
  * cosine top-k       -- the catalog is split into contiguous row shards; every rank computes its
    local top-k with GLOBAL indices, the (dist, idx) lists are all-gathered (k*12 bytes per query
    per rank) and merged in the contract order, so the answer is independent of the shard count.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n`` units for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_gradients(parameters: Sequence[torch.nn.Parameter], group=None, dense_bucket_numel: int = 1 << 22) -> int:
    """Average ``.grad`` over the data-parallel group.  Small tensors are packed into one flat
    bucket (latency-bound ~1 MB message for the dense tower); tensors larger than the bucket
    (embedding tables) are reduced in place.  Returns the number of collectives issued."""
    rank, world = world_info(group)
    if world == 1:
        return 0
    grads = [p.grad for p in parameters if p.grad is not None]
    small = [g for g in grads if g.numel() <= dense_bucket_numel]
    large = [g for g in grads if g.numel() > dense_bucket_numel]
    n_coll = 0
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
    for g in large:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        g.div_(world)
        n_coll += 1
    return n_coll


def gather_topk_and_merge(dist_local: torch.Tensor, idx_local: torch.Tensor,
                          merge_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                          group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's local top-k ([q,k] with global indices) and merge them with
    ``merge_fn(dist_parts [P,q,k], idx_parts [P,q,k])`` (``dcnr_b200.merge_shards`` on the GPU)."""
    rank, world = world_info(group)
    if world == 1:
        return merge_fn(dist_local.unsqueeze(0), idx_local.unsqueeze(0))
    d_parts = [torch.empty_like(dist_local) for _ in range(world)]
    i_parts = [torch.empty_like(idx_local) for _ in range(world)]
    dist.all_gather(d_parts, dist_local.contiguous(), group=group)
    dist.all_gather(i_parts, idx_local.contiguous(), group=group)
    return merge_fn(torch.stack(d_parts), torch.stack(i_parts))


class ShardedNearestNeighbors:
    """The catalog of ``NearestNeighbors`` split over the ranks of ``group`` (configuration 4:
    10 M hotels over 8 B200).  ``fit`` takes THIS rank's contiguous shard and its global row offset."""

    def __init__(self, n_neighbors: int = 16, group=None):
        from .knn import NearestNeighbors
        self.group = group
        self.n_neighbors = n_neighbors
        self._local = None
        self._cls = NearestNeighbors

    def fit_shard(self, shard, index_base: int):
        self._local = self._cls(n_neighbors=self.n_neighbors, metric="cosine", algorithm="brute", index_base=index_base)
        self._local._allow_short = True
        self._local.fit(shard)
        return self

    def kneighbors_tensor(self, Q: torch.Tensor, n_neighbors: Optional[int] = None):
        from .knn import merge_shards
        d, i = self._local.kneighbors_tensor(Q, n_neighbors)
        return gather_topk_and_merge(d, i, merge_shards, self.group)
