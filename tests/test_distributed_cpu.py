"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: request sharding, gradient
all-reduce, and all-gather + merge of sharded top-k lists.  Compute on the CPU side is the oracle
(as the checker); the communication helpers under test are the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcnr_b200 import distributed as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def _grad_job(rank, world):
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(64, 4))]
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = D.allreduce_gradients(ps, dense_bucket_numel=100)       # third tensor (256 elems) goes alone
    return n, [float(p.grad.mean()) for p in ps]


def test_allreduce_gradients_averages_over_ranks():
    out = _run(_grad_job)
    for n, means in out:
        assert n == 2
        assert means == pytest.approx([1.5, 3.0, 4.5])


def _topk_job(rank, world):
    from oracle import knn_oracle
    rng = np.random.default_rng(0)
    E = rng.standard_normal((4000, 16)).astype(np.float32); E[3000] = E[17]
    Q = E[[17, 5]]
    ehat, qhat = knn_oracle.normalize_rows(E), knn_oracle.normalize_rows(Q)
    b, e = D.shard_range(E.shape[0], rank, world)
    d, i = knn_oracle.cosine_topk(ehat[b:e], qhat, 50, idx_base=b)          # per-shard result (oracle as stand-in)

    def merge_fn(dp, ip):
        md, mi = knn_oracle.merge_topk(dp.numpy(), ip.numpy())
        return torch.from_numpy(md), torch.from_numpy(mi)
    md, mi = D.gather_topk_and_merge(torch.from_numpy(d), torch.from_numpy(i), merge_fn)
    fd, fi = knn_oracle.cosine_topk(ehat, qhat, 50)
    return bool((mi.numpy() == fi).all() and (md.numpy() == fd).all())


def test_sharded_topk_gather_and_merge_equals_unsharded():
    assert all(_run(_topk_job))


class _GlooComm:
    """Same interface as distributed.Communicator, over gloo (CPU): the stand-in transport for the host logic."""

    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allgather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous())
        return torch.stack(parts)

    def alltoallv(self, send, send_rows, recv_rows):
        outs = [torch.empty((r,) + tuple(send.shape[1:]), dtype=send.dtype) for r in recv_rows]
        ins = list(torch.split(send.contiguous(), list(send_rows)))
        reqs = []
        for p in range(self.world):
            if p == self.rank:
                outs[p].copy_(ins[p])
            else:
                reqs.append(dist.isend(ins[p], p)); reqs.append(dist.irecv(outs[p], p))
        for r in reqs:
            r.wait()
        return torch.cat(outs)


def _exchange_job(rank, world):
    torch.manual_seed(0)
    full = torch.randn(1001, 8)                                   # odd row count: shards differ in size
    shard = full[rank::world].contiguous()
    g = torch.Generator().manual_seed(7 + rank)
    ids = torch.cat([torch.randint(0, 1001, (300 + 17 * rank,), generator=g), torch.tensor([0, 1000, 1000, 5])])
    rows, plan = D.exchange_lookup(_GlooComm(), shard, ids, gather_rows=lambda s, i: s[i])
    ok_rows = bool(torch.equal(rows, full[ids]))
    owner, local = D.owner_of(ids, world)
    ok_plan = sum(plan["send_rows"]) == ids.numel() and plan["send_rows"][rank] == int((owner == rank).sum())
    ok_local = bool((plan["ids_here"] < shard.shape[0]).all())
    return ok_rows and ok_plan and ok_local


def _exchange_multi_job(rank, world):
    torch.manual_seed(1)
    full = [torch.randn(1001, 8), torch.randn(257, 8)]            # two tables, odd row counts
    shards = [f[rank::world].contiguous() for f in full]
    g = torch.Generator().manual_seed(11 + rank)
    ids = [torch.cat([torch.randint(0, 1001, (200 + 13 * rank,), generator=g), torch.tensor([0, 1000, 1000])]),
           torch.cat([torch.randint(0, 257, (90 + 5 * rank,), generator=g), torch.tensor([256, 0])])]
    rows, plan = D.exchange_lookup_multi(_GlooComm(), shards, ids, gather_rows=lambda s, i: s[i])
    ok = all(bool(torch.equal(r, f[i])) for r, f, i in zip(rows, full, ids))
    ok = ok and sum(plan["send_rows"]) == sum(i.numel() for i in ids)
    return ok


def test_fused_two_table_exchange_returns_the_right_rows():
    assert all(_run(_exchange_multi_job))
    assert all(_run(_exchange_multi_job, world=3))


def test_row_sharded_exchange_returns_the_right_rows_in_batch_order():
    assert all(_run(_exchange_job))
    assert all(_run(_exchange_job, world=3))
