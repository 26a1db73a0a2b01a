"""Multi-GPU plumbing for the DCN-R path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

What shards and how:
  * ranking inference  -- requests are independent: contiguous request ranges per rank, weights
    replicated, NO collective on the data path (results are gathered by the caller);
  * training           -- data parallel: the batch is split, parameters replicated, gradients
    reduced with one flat-bucket all-reduce (dense parameters) plus one all-reduce per embedding
    table.  With a ``Communicator`` attached to the model (``attach``), train-mode BatchNorm
    statistics and the BatchNorm backward sums are taken over the GLOBAL batch inside the library's
    forward / backward calls (NCCL all-gather of a few KB per layer), so an N-rank step equals the
    reference's single-device step on the concatenated batch;
  * row-sharded tables -- ``RowShardedTable`` / ``RowShardedDCN``: user / item tables too large to
    replicate are split by ``row % world``; ids go to the owning rank, rows come back (NCCL
    all-to-all), gradient rows return to the owner, which runs the sorted-segment scatter-add;
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


class Communicator:
    """The library's own NCCL communicator (``dcnr_comm_*`` in include/dcnr.h): created once per
    process from an id that rank 0 generates and ``torch.distributed`` broadcasts.  Its handle is what
    ``dcnr_dims.comm`` carries into the stream-ordered forward / backward calls (SyncBN) and what the
    gradient all-reduce / embedding all-to-all below run on."""

    def __init__(self, group=None, device=None):
        from . import _cabi as C
        import ctypes
        self.rank, self.world = world_info(group)
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            C.check(C.lib().dcnr_comm_unique_id(ident.data_ptr()))
        if self.world > 1:
            on_dev = dist.get_backend(group) == "nccl"
            t = ident.to(self.device) if on_dev else ident
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = t.cpu()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            C.check(C.lib().dcnr_comm_create(ident.data_ptr(), self.rank, self.world, ctypes.byref(self._h)))

    @property
    def handle(self):
        return self._h

    @property
    def uses_peer_memory(self) -> bool:
        """True when the small BatchNorm exchanges run as peer-to-peer kernels over NVLink instead of NCCL all-gathers."""
        from . import _cabi as C
        return bool(C.lib().dcnr_comm_uses_peer_memory(self._h))

    def set_peer_memory(self, enable: bool) -> None:
        """Park (False) or restore (True) the peer-memory exchanges; same value on every rank, between steps."""
        from . import _cabi as C
        C.check(C.lib().dcnr_comm_set_peer_memory(self._h, int(bool(enable))))

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        from . import _cabi as C
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        C.check(C.lib().dcnr_comm_allreduce_f32(self._h, C.ptr(t), t.numel(), C.stream()))
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        """[world, *t.shape] with rank r's tensor at index r."""
        from . import _cabi as C
        t = t.contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        C.check(C.lib().dcnr_comm_allgather(self._h, C.ptr(t), C.ptr(out), t.numel() * t.element_size(), C.stream()))
        return out

    def alltoallv(self, send: torch.Tensor, send_rows: Sequence[int], recv_rows: Sequence[int]) -> torch.Tensor:
        """``send`` holds the rows for rank 0, then rank 1, ... (``send_rows[p]`` each); returns the rows received
        from rank 0, 1, ... (``recv_rows[p]`` each).  Row = everything but the leading dimension."""
        from . import _cabi as C
        import ctypes
        send = send.contiguous()
        row_bytes = send.element_size()
        for dim in send.shape[1:]:
            row_bytes *= int(dim)
        recv = torch.empty((int(sum(recv_rows)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        arr = lambda v: (ctypes.c_int64 * self.world)(*[int(x) for x in v])
        offs = lambda rows: [int(sum(rows[:p])) * row_bytes for p in range(self.world)]
        C.check(C.lib().dcnr_comm_alltoallv(self._h, C.ptr(send), arr([r * row_bytes for r in send_rows]), arr(offs(send_rows)),
                                            C.ptr(recv), arr([r * row_bytes for r in recv_rows]), arr(offs(recv_rows)),
                                            C.stream()))
        return recv

    def close(self):
        from . import _cabi as C
        if self._h:
            C.lib().dcnr_comm_destroy(self._h)
            self._h = None


def attach(model, comm: Optional[Communicator], batch_capacity: Optional[int] = None):
    """Make ``model`` (a ``DCN_RecSys``) use global-batch BatchNorm statistics over ``comm`` in train().

    ``batch_capacity``: the largest LOCAL batch any rank will feed in one step.  With it the ranks may hold different
    batch sizes (the ragged last batch of an epoch): the sparse table-gradient exchange pads every rank to the capacity.
    ``None`` declares that all ranks always hold the same number of rows (a mismatch then shows up as an NCCL error)."""
    model._comm = comm
    model._dp_batch_cap = int(batch_capacity) if batch_capacity else 0
    return model


def allreduce_gradients(parameters: Sequence[torch.nn.Parameter], group=None, dense_bucket_numel: int = 1 << 22,
                        comm: Optional[Communicator] = None, average: bool = True) -> int:
    """Reduce ``.grad`` over the data-parallel group (mean by default; ``average=False`` sums, for losses
    already scaled by the global batch).  Small tensors are packed into one flat bucket (latency-bound
    ~1 MB message for the dense tower); tensors larger than the bucket (embedding tables) are reduced in
    place.  With ``comm`` the library's communicator is used, else ``torch.distributed``.  Returns the
    number of collectives issued."""
    rank, world = (comm.rank, comm.world) if comm is not None else world_info(group)
    if world == 1:
        return 0
    grads = [p.grad for p in parameters if p.grad is not None]
    small = [g for g in grads if g.numel() <= dense_bucket_numel]
    large = [g for g in grads if g.numel() > dense_bucket_numel]
    n_coll = 0

    def reduce_(t):
        if comm is not None:
            comm.allreduce_(t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if average:
            t.div_(world)
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        reduce_(flat)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
    for g in large:
        reduce_(g)
        n_coll += 1
    return n_coll


# ---------------------------------------------------------------------------------------------------
# Row-sharded embedding tables (BASELINE configs[4]: 100 M-row user / item tables over 8 B200)
# ---------------------------------------------------------------------------------------------------
def owner_of(ids: torch.Tensor, world: int):
    """Row r of a sharded table lives on rank r % world at local row r // world (balances Zipf ids)."""
    return ids % world, ids // world


def _gather_rows_device(shard: torch.Tensor, ids_here: torch.Tensor) -> torch.Tensor:
    from . import _cabi as C
    rows_here = torch.empty((ids_here.numel(), shard.shape[1]), dtype=torch.float32, device=shard.device)
    C.check(C.lib().dcnr_gather_rows(C.ptr(shard), shard.shape[0], shard.shape[1], C.ptr(ids_here), ids_here.numel(),
                                     C.ptr(rows_here), C.stream()))
    return rows_here


def exchange_lookup(comm, shard: torch.Tensor, ids: torch.Tensor, gather_rows=_gather_rows_device):
    """Forward half of the row-sharded exchange (host logic; ``comm`` needs rank / world / allgather / alltoallv,
    ``gather_rows(shard, local_ids)`` is the owner-side lookup -- the library kernel in the product, injectable so
    the world_size-2 gloo tests can run this logic on CPU).  Returns (rows in batch order, plan for backward)."""
    world, rank = comm.world, comm.rank
    owner, local = owner_of(ids, world)
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world)
    matrix = comm.allgather(counts).cpu() if world > 1 else counts.reshape(1, 1).cpu()   # one host sync: message sizes
    send_rows = [int(v) for v in matrix[rank]]               # ids this rank asks rank p for
    recv_rows = [int(v) for v in matrix[:, rank]]            # ids rank p asks this rank for
    ids_sorted = local[order].contiguous()
    ids_here = comm.alltoallv(ids_sorted, send_rows, recv_rows) if world > 1 else ids_sorted
    rows_here = gather_rows(shard, ids_here)
    rows_sorted = comm.alltoallv(rows_here, recv_rows, send_rows) if world > 1 else rows_here
    rows = torch.empty_like(rows_sorted)
    rows[order] = rows_sorted                                # back to batch order
    return rows, dict(order=order, ids_here=ids_here, send_rows=send_rows, recv_rows=recv_rows)


def exchange_lookup_multi(comm, shards, ids_list, gather_rows=_gather_rows_device):
    """``exchange_lookup`` for several tables of the same width in ONE exchange: one owner sort, one count all-gather (one
    host sync), one all-to-all of ids and one of rows for all tables together -- the user and the item lookup of a step cost
    the collectives of one.  Segment layout of every message: for peer p the ids of table 0, then table 1, ...
    Returns ([rows of table t in batch order], plan)."""
    world, rank, T = comm.world, comm.rank, len(shards)
    dev = ids_list[0].device
    if world == 1:                                            # every row is local: no sort, no exchange
        sizes = [ids.numel() for ids in ids_list]
        ids_here = torch.cat(ids_list)
        bounds = [0]
        for n in sizes:
            bounds.append(bounds[-1] + n)
        pos = [torch.arange(bounds[t], bounds[t + 1], device=dev) for t in range(T)]
        rows = [gather_rows(shards[t], ids_list[t].contiguous()) for t in range(T)]
        return rows, dict(order=None, ids_here=ids_here, pos=pos, send_rows=[ids_here.numel()], recv_rows=[ids_here.numel()],
                          sizes=sizes)
    keys, locals_ = [], []
    for t, ids in enumerate(ids_list):
        owner, local = owner_of(ids, world)
        keys.append(owner * T + t)
        locals_.append(local)
    key, local = torch.cat(keys), torch.cat(locals_)
    order = torch.sort(key, stable=True).indices
    counts = torch.bincount(key, minlength=world * T)
    matrix = (comm.allgather(counts).cpu() if world > 1 else counts.reshape(1, -1).cpu()).reshape(world, world, T)
    send_seg = matrix[rank]                                   # [peer, table]: ids this rank asks peer p for
    recv_seg = matrix[:, rank, :]                             # [peer, table]: ids peer p asks this rank for
    send_rows = [int(v) for v in send_seg.sum(1)]
    recv_rows = [int(v) for v in recv_seg.sum(1)]
    ids_sorted = local[order].contiguous()
    ids_here = comm.alltoallv(ids_sorted, send_rows, recv_rows) if world > 1 else ids_sorted
    # which table each received id belongs to: segments (peer 0: t0, t1, ...), (peer 1: ...)
    # (all sizes are known on the host from `matrix`: no further sync)
    n_here = int(recv_seg.sum())
    table_here = torch.repeat_interleave(torch.arange(T, device=dev).repeat(world), recv_seg.reshape(-1).to(dev),
                                         output_size=n_here)
    pos = list(torch.split(torch.sort(table_here, stable=True).indices, [int(v) for v in recv_seg.sum(0)]))
    rows_here = torch.empty((ids_here.numel(), shards[0].shape[1]), dtype=shards[0].dtype, device=dev)
    for t in range(T):
        rows_here[pos[t]] = gather_rows(shards[t], ids_here[pos[t]].contiguous())
    rows_sorted = comm.alltoallv(rows_here, recv_rows, send_rows) if world > 1 else rows_here
    rows = torch.empty_like(rows_sorted)
    rows[order] = rows_sorted                                 # back to (table, batch) order
    sizes = [ids.numel() for ids in ids_list]
    return list(torch.split(rows, sizes)), dict(order=order, ids_here=ids_here, pos=pos, send_rows=send_rows, recv_rows=recv_rows,
                                                sizes=sizes)


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table: "RowShardedTable", shard: torch.Tensor, ids: torch.Tensor):
        rows, plan = exchange_lookup(table.comm, shard, ids)
        rank = table.comm.rank
        ctx.table, ctx.order, ctx.ids_here = table, plan["order"], plan["ids_here"]
        ctx.send_rows, ctx.recv_rows, ctx.shard_shape = plan["send_rows"], plan["recv_rows"], tuple(shard.shape)
        table.last_exchange_bytes = ((sum(ctx.send_rows) - ctx.send_rows[rank]) * 8 +
                                     (sum(ctx.recv_rows) - ctx.recv_rows[rank]) * shard.shape[1] * 4)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        from . import _cabi as C
        from .functional import _scratch
        table, comm = ctx.table, ctx.table.comm
        g_sorted = grad_rows.contiguous()[ctx.order].contiguous()
        g_here = comm.alltoallv(g_sorted, ctx.send_rows, ctx.recv_rows) if comm.world > 1 else g_sorted
        grad_shard = torch.empty(ctx.shard_shape, dtype=torch.float32, device=grad_rows.device)
        n = ctx.ids_here.numel()
        ws = _scratch(table.scatter_scratch_bytes(n), grad_rows.device)
        C.check(C.lib().dcnr_scatter_rows(C.ptr(ctx.ids_here), n, ctx.shard_shape[0], ctx.shard_shape[1], C.ptr(g_here),
                                          ctx.shard_shape[1], C.ptr(grad_shard), C.ptr(ws), ws.numel(), C.stream()))
        return None, grad_shard, None


class _LookupMultiFn(torch.autograd.Function):
    """Two (or more) row-sharded tables through one exchange: inputs (tables tuple, shard_0, ids_0, shard_1, ids_1, ...)."""

    @staticmethod
    def forward(ctx, tables, *args):
        shards, ids_list = list(args[0::2]), list(args[1::2])
        comm = tables[0].comm
        rows, plan = exchange_lookup_multi(comm, shards, ids_list)
        ctx.tables, ctx.plan, ctx.shapes = tables, plan, [tuple(s.shape) for s in shards]
        dim, rank = shards[0].shape[1], comm.rank
        moved = ((sum(plan["send_rows"]) - plan["send_rows"][rank]) * 8 + (sum(plan["recv_rows"]) - plan["recv_rows"][rank]) * dim * 4)
        for t in tables:
            t.last_exchange_bytes = moved // len(tables)
        return tuple(rows)

    @staticmethod
    def backward(ctx, *grad_rows):
        from . import _cabi as C
        from .functional import _scratch
        plan, comm = ctx.plan, ctx.tables[0].comm
        g = torch.cat([gr.contiguous() for gr in grad_rows])
        if plan["order"] is None:                             # single rank
            g_here = g
        else:
            g_sorted = g[plan["order"]].contiguous()
            g_here = comm.alltoallv(g_sorted, plan["send_rows"], plan["recv_rows"])
        out = [None]
        for t, table in enumerate(ctx.tables):
            shape = ctx.shapes[t]
            ids_t = plan["ids_here"][plan["pos"][t]].contiguous()
            g_t = g_here[plan["pos"][t]].contiguous()
            grad_shard = torch.empty(shape, dtype=torch.float32, device=g.device)
            n = ids_t.numel()
            ws = _scratch(table.scatter_scratch_bytes(n), g.device)
            C.check(C.lib().dcnr_scatter_rows(C.ptr(ids_t), n, shape[0], shape[1], C.ptr(g_t), shape[1], C.ptr(grad_shard),
                                              C.ptr(ws), ws.numel(), C.stream()))
            out += [grad_shard, None]
        return tuple(out)


def lookup_tables(tables, ids_list):
    """Rows of several ``RowShardedTable`` s (same width, same communicator) through one fused exchange."""
    args = []
    for table, ids in zip(tables, ids_list):
        args += [table.weight, ids.reshape(-1).to(torch.int64).contiguous()]
    return _LookupMultiFn.apply(tuple(tables), *args)


class RowShardedTable(torch.nn.Module):
    """An ``nn.Embedding(n_rows, dim)`` split over the ranks of ``comm`` by ``row % world``.  ``forward(ids)``
    returns the rows in batch order (ids -> owners, owners gather with ``dcnr_gather_rows``, rows -> back);
    backward sends the gradient rows to the owners, which run the deterministic sorted-segment scatter-add
    (``dcnr_scatter_rows``) into the dense gradient of their shard."""

    def __init__(self, n_rows: int, dim: int, comm: Communicator, device=None, init_std: float = 1.0):
        super().__init__()
        self.comm, self.n_rows, self.dim = comm, int(n_rows), int(dim)
        local = (self.n_rows - comm.rank + comm.world - 1) // comm.world
        self.weight = torch.nn.Parameter(torch.empty((max(local, 1), dim), dtype=torch.float32, device=device))
        torch.nn.init.normal_(self.weight, std=init_std)
        self.last_exchange_bytes = 0

    def scatter_scratch_bytes(self, n: int) -> int:
        from . import _cabi as C
        d = C.Dims()
        d.emb_dim, d.hidden, d.in_dim, d.in_dim_pad = self.dim, 4, 2 * self.dim, C.pad_dim(2 * self.dim)
        return int(C.lib().dcnr_workspace_bytes(d, max(n, 1), 2))

    def load_full(self, full: torch.Tensor):
        """Take this rank's rows out of the full [n_rows, dim] table (tests / checkpoint loading)."""
        with torch.no_grad():
            self.weight.copy_(full[self.comm.rank::self.comm.world].to(self.weight.device))
        return self

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return _LookupFn.apply(self, self.weight, ids.reshape(-1).to(torch.int64).contiguous())


class RowShardedDCN(torch.nn.Module):
    """DCN-R with row-sharded user / item tables and a replicated dense part (configs[4]).  ``core`` is a
    ``DCN_RecSys`` whose own user / item tables are unused 1-row placeholders: the rows fetched by the two
    ``RowShardedTable`` exchanges are handed to the library as per-sample tables (id = batch position), so the
    fused gather + cross kernel, the dense tower and the backward are the unmodified single-GPU path."""

    def __init__(self, n_users: int, n_items: int, cat_dims, n_num_features: int, params, comm: Communicator,
                 precision: Optional[str] = None, device=None):
        super().__init__()
        from .model import DCN_RecSys
        self.comm = comm
        self.core = DCN_RecSys(1, 1, cat_dims, n_num_features, params, precision)
        self.user_table = RowShardedTable(n_users, params["emb_dim"], comm, device)
        self.item_table = RowShardedTable(n_items, params["emb_dim"], comm, device)
        if device is not None:
            self.core.to(device)
        attach(self.core, comm)

    def forward(self, user_ids, item_ids, cat_features, num_features):
        user_rows, item_rows = lookup_tables([self.user_table, self.item_table], [user_ids, item_ids])      # one exchange
        return self.core.forward_rows(user_rows, item_rows, cat_features, num_features)

    def dense_parameters(self):
        skip = {id(self.core.user_embedding.weight), id(self.core.item_embedding.weight)}
        return [p for p in self.core.parameters() if id(p) not in skip]


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
