#!/usr/bin/env python
"""bench.py -- headline benchmark of the DCN-R hot path on B200.

Workload (BASELINE.json configs[1]): DCN-R ranking inference, 1 user x 500 candidates per request,
65 536 synthetic requests batched = 32 768 000 candidate rows per step, model P0 (emb 16, hidden
256, 3 cross layers, 2 ResBlocks), tables 1 M users x 100 K hotels, eval mode.
A "step" is one pass of the ranking forward over all rows.

  value     candidates/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same through serving.RankingEngine with HOST (pinned) buffers: H2D of the inputs
            and D2H of the scores inside the timed region
  roofline  the dominant kernel (the H x H dense layer of the ResBlocks) timed live
  cpu_baseline / --impl reference: the oracle port of the reference's CPU PyTorch path
            (oracle/dcnr_oracle.py, same ATen ops) on the box's host cores, bounded sample
Extra (not part of the contract line's headline): "train" = configs[2] training step (data parallel with
global-batch BatchNorm + NCCL gradient all-reduce when N > 1), "similarity" = configs[3] cosine top-k,
"sharded" = configs[4] row-sharded 100 M-row tables with the all-to-all exchange, "kernels" = the HBM-bound
kernels (gather, scatter, top-k scan) against their algorithmic bytes -- each with its own unit.

N > 1 (torchrun): requests are sharded, every rank scores its own 32 768 000 rows (weak scaling),
no collective on the data path; value = all ranks' rows / max-over-ranks time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P0 = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.6)
N_USERS, N_ITEMS = 1_000_000, 100_000
CAT_DIMS = {"city": 100, "hotel_type": 6}
N_NUM = 11
REQUESTS, CANDIDATES = 65_536, 500
FLOP_PER_ROW = 2 * 57 * 256 + 2 * 2 * 2 * 256 * 256 + 2 * 313          # 554 098


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p["bf16_tflops_sustained"], tensor_burst=p["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]            # upper half = samples under load
            out["sm_mhz"] = statistics.median(busy)
        out["reasons"], out["samples"] = sorted(reasons), len(sm)
        return out


def make_state(seed=42):
    """P0 parameters, 'trained-like' (SURVEY.md 8d): embeddings x0.1, randomised BN statistics."""
    from oracle import dcnr_oracle as orc          # only used as an initialiser of synthetic weights
    return orc.make_state(N_USERS, N_ITEMS, CAT_DIMS, N_NUM, P0, seed=seed, emb_scale=0.1, randomize_bn=True)


def synth_state_device(dev, seed=42):
    """Same shapes as make_state but generated on the device (1 M x 16 tables)."""
    import dcnr_b200
    torch.manual_seed(seed)
    m = dcnr_b200.DCN_RecSys(N_USERS, N_ITEMS, CAT_DIMS, N_NUM, P0)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for e in [m.user_embedding, m.item_embedding, *m.cat_embeddings]:
            e.weight.mul_(0.1)
        for blk in m.res_blocks:
            for bn in (blk.bn1, blk.bn2):
                bn.weight.copy_(0.5 + torch.rand(bn.weight.shape, generator=g))
                bn.bias.copy_(torch.randn(bn.bias.shape, generator=g) * 0.2)
                bn.running_mean.copy_(torch.randn(bn.bias.shape, generator=g) * 0.3)
                bn.running_var.copy_(0.5 + torch.rand(bn.bias.shape, generator=g))
    return m.to(dev)


def synth_requests(n_req, n_cand, seed, device):
    """Ranking inputs in the hackathon_augmented_data.csv tensor schema (main.py:215-230): one user
    index repeated per request, item ids, (city, hotel_type) as functions of the hotel, 11 scaled numerics."""
    g = torch.Generator(device=device).manual_seed(seed)
    users = torch.randint(0, N_USERS, (n_req,), generator=g, device=device)
    user_ids = users.repeat_interleave(n_cand)
    item_ids = torch.randint(0, N_ITEMS, (n_req * n_cand,), generator=g, device=device)
    city_of = torch.randint(0, CAT_DIMS["city"], (N_ITEMS,), generator=g, device=device)
    type_of = torch.randint(0, CAT_DIMS["hotel_type"], (N_ITEMS,), generator=g, device=device)
    cat = torch.stack([city_of[item_ids], type_of[item_ids]], dim=1).contiguous()
    num = torch.rand((n_req * n_cand, N_NUM), generator=g, device=device)
    return user_ids, item_ids, cat, num


def time_steps(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for s in range(steps):
        fn()
        ev[s + 1].record()
    torch.cuda.synchronize()
    barrier()
    return ev[0].elapsed_time(ev[-1]) / 1e3       # seconds for exactly `steps` steps


def reference_arm(args):
    """The reference's CPU PyTorch path (oracle port: same ATen ops as main.DCN_RecSys) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import dcnr_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    state = make_state()
    n_req = 128                                                     # bounded sample per step: 64 000 rows
    u, i, c, x = synth_requests(n_req, CANDIDATES, 1234, "cpu")
    rows = u.numel()

    def step():
        with torch.no_grad():
            orc.forward(state, u, i, c, x, training=False)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = f"{n_req} requests x {CANDIDATES} candidates = {rows} rows per step (of {REQUESTS * CANDIDATES})"
    _emit({
        "impl": "reference", "metric": "ranking_candidates_per_s", "value": value, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DCN-R P0 ranking inference, 1 user x 500 candidates/request (BASELINE configs[1])",
                   "tables": f"{N_USERS} users x {N_ITEMS} hotels", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def _emit(obj):
    """The contract line goes to the process's ORIGINAL stdout; everything else that writes to fd 1 while the bench
    runs (NCCL prints its version banner there) has been redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dcnr_b200", choices=["dcnr_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DCNR_PRECISION", "tf32x3"),
                    help="dense-layer arithmetic: tf32x3 (tcgen05, fp32-parity split; default), fp32 (CUDA cores), tf32")
    ap.add_argument("--requests", type=int, default=REQUESTS)
    ap.add_argument("--skip-extras", action="store_true", help="skip the train / similarity side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import dcnr_b200
    from dcnr_b200 import _cabi as C
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (the product has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pk = peaks()
    model = synth_state_device(dev).eval()
    model.precision = args.precision
    rows = args.requests * CANDIDATES
    u, i, c, x = synth_requests(args.requests, CANDIDATES, 1234 + rank, dev)
    out_dev = {}

    def step_resident():
        with torch.no_grad():
            out_dev["logits"] = model(u, i, c, x)

    # ---- value: inputs resident in HBM ----------------------------------------------------------
    C.launch_count(reset=True)
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    C.launch_count(reset=True)
    secs = time_steps(step_resident, args.steps, 0, barrier)
    launches = C.launch_count()
    clocks = sampler.stop()
    secs = max_over_ranks(secs)
    value = world * rows * args.steps / secs
    checksum = float(out_dev["logits"].double().sum())

    # ---- e2e: host buffers through the public serving call ----------------------------------------
    hu, hi, hc, hx = (t.cpu().pin_memory() for t in (u, i, c, x))
    engine = dcnr_b200.serving.RankingEngine(model, chunk_rows=1 << 20)
    hout = torch.empty(rows, dtype=torch.float32, pin_memory=True)
    e2e_steps = max(2, min(args.steps, 3))
    e2e_secs = max_over_ranks(time_steps(lambda: engine.score(hu, hi, hc, hx, out=hout), e2e_steps, 1, barrier))
    e2e_value = world * rows * e2e_steps / e2e_secs
    assert abs(float(hout.double().sum()) - checksum) <= 1e-6 * max(1.0, abs(checksum)) * 10, "e2e result differs"

    # ---- roofline of the dominant kernel: the H x H dense layer (4 of the 5 GEMMs, 95 % of the FLOPs) ----
    M, H = min(rows, 1 << 20), P0["hidden_dim"]
    a = torch.randn(M, H, device=dev); w = torch.randn(H, H, device=dev) / 16
    sc = torch.rand(H, device=dev); sh = torch.rand(H, device=dev)
    prec = C.PRECISIONS[args.precision]
    k_secs = time_steps(lambda: dcnr_b200.functional.linear_forward_raw(a, w, sh, sc, a, True, prec), 10, 3, lambda: None)
    k_flops = 2.0 * M * H * H
    isolated = k_flops * 10 / k_secs / 1e12
    del a, w
    # live: a repeat of the timed steps with a CUDA event pair (on the launching stream) around EVERY dense-layer GEMM launch
    C.gemm_timing_begin()
    t_rep = time_steps(step_resident, args.steps, 0, lambda: None)
    g_ms, g_n, g_fl = C.gemm_timing_end()
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    kname = "k_sgemm<true,true> (fp32 CUDA-core)" if args.precision == "fp32" else f"k_gemm_tc, tcgen05 ({args.precision})"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": achieved / pk["tensor"], "traffic": None, "kernel": kname,
                "launches": g_n, "avg_launch_us": g_ms * 1e3 / max(g_n, 1), "share_of_step": g_ms * 1e-3 / t_rep,
                "how": f"CUDA event pairs on the launching stream around every dense-layer GEMM launch over a repeat of the "
                       f"{args.steps} timed steps ({g_n} launches: per 1 Mi-row chunk one 256x64 initial layer + four 256x256 "
                       f"layers with the folded-BN / residual / ReLU epilogues); achieved = sum of algorithmic flops (2mnk) / "
                       f"sum of launch durations; peak = bf16 sustained ({pk['src']})",
                "isolated_layer_tflops": isolated,
                "isolated_how": f"dcnr_linear_fwd {M}x{H}x{H} + scale/shift/residual/relu alone, 10 launches after 3 warm-ups",
                "step_algorithmic_tflops": FLOP_PER_ROW * rows * args.steps / secs / 1e12}

    result = {
        "metric": "ranking_candidates_per_s", "value": value, "unit": "candidates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-parity split, fp32 accumulate)", "tf32": "tf32", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "DCN-R P0 ranking inference, 1 user x 500 candidates/request, "
                               f"{args.requests} requests/step per GPU (BASELINE configs[1])",
                   "rows_per_step_per_gpu": rows, "tables": f"{N_USERS} users x {N_ITEMS} hotels",
                   "model": "emb16 hidden256 cross3 res2 (P0)", "l2": "inputs 2.5 GB per step > 126 MB L2, no flush needed",
                   "parallelism": f"requests sharded x{world}, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": rows * engine.bytes_per_row_h2d,
                "d2h_bytes_per_step": rows * engine.bytes_per_row_d2h, "steps": e2e_steps,
                "ms_per_step": e2e_secs / e2e_steps * 1e3},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "checksum": checksum,
    }

    traffic_file = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(traffic_file):                        # dram bytes per launch of the same kernel from the ncu --set full capture
        t = json.load(open(traffic_file)).get(args.precision)
        if t:
            roofline["traffic"] = t["dram_bytes_per_launch"]
            roofline["traffic_note"] = t["note"]
    if not args.skip_extras:
        comm = None
        if world > 1:
            comm = dcnr_b200.distributed.Communicator()
        result["train"] = bench_train(model, dev, world, rank, comm, barrier, max_over_ranks)
        if world > 1:      # the same data-parallel step with 65 536 rows PER GPU: what the sync costs when every GPU has real work
            tw = bench_train(model, dev, world, rank, comm, barrier, max_over_ranks, global_batch=65_536 * world, steps=5)
            tw["scaling"] = "weak"
            result["train_weak"] = tw
        del model
        torch.cuda.empty_cache()
        result["sharded"] = bench_sharded(dev, world, rank, comm, barrier, max_over_ranks, args.precision)
        torch.cuda.empty_cache()
        if world == 1:
            result["mmr"] = bench_mmr(dev)
            result["similarity"] = bench_similarity(dev, pk)
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import kernel_probe
            result["kernels"] = {k: {kk: vv for kk, vv in v.items() if kk != "note"} for k, v in
                                 kernel_probe.probe(65_536, pk["hbm"]).items()}
            torch.cuda.empty_cache()
            result["kernels_4m_rows"] = {k: {kk: vv for kk, vv in v.items() if kk != "note"} for k, v in
                                         kernel_probe.probe(1 << 22, pk["hbm"]).items()}
            result["kernels_4m_rows"]["how"] = "the same operator calls at B = 4 194 304 rows (the bandwidth regime: every tensor > L2)"
            result["kernels"]["how"] = ("C-ABI operator calls at the configs[2] shape (B = 65536, tables 1M x 100K), CUDA events, "
                                        "algorithmic bytes per row from SURVEY.md 8d; hbm_frac = GB/s / measured copy bandwidth")
        if comm is not None:
            comm.close()
    if rank == 0 and world == 1:
        # (--skip-extras is the profiling form of the command: keep torch's own kernels out of its launch list)
        result["cpu_baseline"] = cpu_baseline(dev=None if args.skip_extras else dev)
    if rank == 0:
        _emit(result)
    if dist is not None:
        dist.destroy_process_group()


def bench_train(model, dev, world, rank, comm, barrier, max_over_ranks, global_batch=65_536, steps=10):
    """BASELINE configs[2]: P0 training step (fwd + BCE + bwd + gradient all-reduce), global batch 65 536
    split over the ranks (strong scaling), 1 M x 100 K tables, dropout 0.6, dense embedding grads.  With N > 1
    the BatchNorm statistics cover the global batch (library communicator), the loss gradient is scaled by the
    global batch and the gradients are SUMMED over ranks: the step is the single-device step on 65 536 rows."""
    import dcnr_b200
    from dcnr_b200 import _cabi as C
    from dcnr_b200.distributed import allreduce_gradients, attach
    B = global_batch // world
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    u = torch.randint(0, N_USERS, (B,), generator=g, device=dev)
    i = torch.randint(0, N_ITEMS, (B,), generator=g, device=dev)
    c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in CAT_DIMS.values()], 1)
    x = torch.rand((B, N_NUM), generator=g, device=dev)
    y = (torch.rand(B, generator=g, device=dev) < 0.3).float()
    model.train()
    attach(model, comm)
    params = list(model.parameters())

    def step():
        for p in params:
            p.grad = None
        logits = model(u, i, c, x)
        _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
        if world > 1:
            dl = dl / world                                   # mean over the GLOBAL batch
        logits.backward(gradient=dl)
        allreduce_gradients(model.parameters_to_allreduce(), comm=comm, average=False)
    C.launch_count(reset=True)
    step(); torch.cuda.synchronize()
    per_step = C.launch_count()
    secs = max_over_ranks(time_steps(step, steps, 3, barrier))
    # the same step captured once as a CUDA graph (dcnr_b200.training.GraphedTrainStep) and replayed
    graphed = None
    try:
        gs = dcnr_b200.training.GraphedTrainStep(model, B, comm)
        gs.load(u, i, c, x, y)
        gs.capture()
        gsecs = max_over_ranks(time_steps(lambda: gs(), 2 * steps, 3, barrier))
        graphed = {"value": global_batch * 2 * steps / gsecs, "unit": "samples/s", "ms_per_step": gsecs / (2 * steps) * 1e3,
                   "loss": float(gs.loss)}
        del gs
    except Exception as e:                                      # the eager number above stands on its own
        graphed = {"error": str(e)[:300]}
    model._dropout_step = None
    model.eval()
    attach(model, None)
    flops = 3 * FLOP_PER_ROW * global_batch
    return {"metric": "train_samples_per_s", "value": global_batch * steps / secs, "unit": "samples/s",
            "global_batch": global_batch, "per_gpu_batch": B, "ms_per_step": secs / steps * 1e3, "scaling": "strong",
            "algorithmic_tflops": flops * steps / secs / 1e12, "cuda_graph": graphed,
            "includes": "forward + BCE + backward + NCCL all-reduce of the dense gradients; when N > 1 also global-batch "
                        "BatchNorm and the all-gather of (id, gradient row) pairs that builds the table gradients "
                        "(optimizer excluded)",
            "gpu_launches_per_step": per_step}


def bench_sharded(dev, world, rank, comm, barrier, max_over_ranks, precision, global_batch=262_144, rows=100_000_000,
                  steps=5):
    """BASELINE configs[4]: user / item tables of 100 M rows x 16 (6.4 GB each) row-sharded over the ranks by
    row % N, global batch 262 144 split over the ranks, forward + backward with the NCCL all-to-all exchange of
    ids / rows / gradient rows and the owner-side sorted-segment scatter-add; dense part data parallel."""
    import dcnr_b200
    from dcnr_b200.distributed import Communicator, RowShardedDCN, allreduce_gradients
    own = comm is None
    if own:
        comm = Communicator()                                 # world 1: no NCCL traffic, same code path
    B = global_batch // world
    model = RowShardedDCN(rows, rows, CAT_DIMS, N_NUM, P0, comm, precision=precision, device=dev).train()
    with torch.no_grad():
        model.user_table.weight.mul_(0.1); model.item_table.weight.mul_(0.1)
    g = torch.Generator(device=dev).manual_seed(321 + rank)
    u = torch.randint(0, rows, (B,), generator=g, device=dev)
    i = torch.randint(0, rows, (B,), generator=g, device=dev)
    c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in CAT_DIMS.values()], 1)
    x = torch.rand((B, N_NUM), generator=g, device=dev)
    y = (torch.rand(B, generator=g, device=dev) < 0.3).float()
    dense = model.dense_parameters()
    every = list(model.parameters())

    def step():
        for p in every:
            p.grad = None
        logits = model(u, i, c, x)
        _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
        logits.backward(gradient=dl / world)
        allreduce_gradients(dense, comm=comm, average=False)
    secs = max_over_ranks(time_steps(step, steps, 2, barrier))
    exch = model.user_table.last_exchange_bytes + model.item_table.last_exchange_bytes
    out = {"metric": "sharded_train_samples_per_s", "value": global_batch * steps / secs, "unit": "samples/s",
           "global_batch": global_batch, "per_gpu_batch": B, "table_rows": rows, "rows_per_rank": model.user_table.weight.shape[0],
           "ms_per_step": secs / steps * 1e3, "scaling": "strong",
           "alltoall_bytes_per_rank_fwd": exch, "includes": "ids/rows all-to-all fwd, gradient rows all-to-all + owner-side "
           "scatter-add bwd (dense shard gradients, zero-filled every step), dense all-reduce; optimizer excluded"}
    del model
    if own:
        comm.close()
    return out


def bench_similarity(dev, pk, n=10_000_000, d=16, k=201):
    """BASELINE configs[3] on one GPU: cosine top-201 over a 10 M x 16 catalog (640 MB scan per query batch).
    hbm_frac = one catalog read / whole-call time (normalise + scan + merge), replayed from a CUDA graph."""
    import dcnr_b200
    g = torch.Generator(device=dev).manual_seed(7)
    E = torch.randn(n, d, device=dev, generator=g)
    model = dcnr_b200.NearestNeighbors().fit(E)
    out = {}
    for nq in (1, 8, 32):
        Q = E[torch.randint(0, n, (nq,), device=dev, generator=g)].contiguous()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            model.kneighbors_tensor(Q, k)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                model.kneighbors_tensor(Q, k)
        torch.cuda.synchronize()
        secs = time_steps(gr.replay, 10, 3, lambda: None) / 10
        passes = 1 if nq == 1 else (nq + 7) // 8
        out[f"q{nq}"] = {"ms": secs * 1e3, "pairs_per_s": nq * n / secs, "catalog_passes": passes,
                         "hbm_frac": (n * d * 4) / secs / 1e9 / pk["hbm"],
                         "hbm_frac_per_pass": passes * (n * d * 4) / secs / 1e9 / pk["hbm"]}
    return {"metric": "similarity_query_candidate_pairs_per_s", "catalog": f"{n} x {d} f32", "k": k, **out}


def bench_mmr(dev, n_req=4096, n_cand=CANDIDATES):
    """the greedy MMR re-rank (main.py:327-332, top_k 20, lambda 0.7) of n_req requests x 500 scored
    candidates in one launch (one CTA per request), next to the oracle's C restatement on one host core."""
    import numpy as np
    from dcnr_b200 import _cabi as C
    g = torch.Generator(device=dev).manual_seed(11)
    emb = torch.randn(N_ITEMS, P0["emb_dim"], generator=g, device=dev)
    idx = torch.randint(0, N_ITEMS, (n_req * n_cand,), generator=g, device=dev)
    scores = torch.sort(torch.randn(n_req, n_cand, generator=g, device=dev), dim=1, descending=True).values.reshape(-1).contiguous()
    offsets = torch.arange(0, (n_req + 1) * n_cand, n_cand, dtype=torch.int32, device=dev)
    order = torch.empty((n_req, 20), dtype=torch.int32, device=dev); count = torch.empty(n_req, dtype=torch.int32, device=dev)

    def run():
        C.check(C.lib().dcnr_mmr_rerank(C.ptr(emb), N_ITEMS, P0["emb_dim"], C.ptr(scores), C.ptr(idx), C.ptr(offsets), n_req, 0.7, 20,
                                        n_cand, C.ptr(order), C.ptr(count), C.stream()))
    secs = time_steps(run, 5, 2, lambda: None) / 5
    out = {"metric": "mmr_requests_per_s", "value": n_req / secs, "unit": "requests/s", "requests": n_req, "candidates": n_cand,
           "ms": secs * 1e3}
    try:
        from oracle import mmr_oracle
        E = emb.cpu().numpy(); sc = scores[: 16 * n_cand].cpu().numpy().reshape(16, n_cand); ix = idx[: 16 * n_cand].cpu().numpy().reshape(16, n_cand)
        t0 = time.perf_counter()
        ref = [mmr_oracle.mmr_rerank(E, sc[r], ix[r], 0.7, 20) for r in range(16)]
        dt = time.perf_counter() - t0
        out["cpu_oracle_requests_per_s"] = 16 / dt
        out["matches_oracle"] = bool(all(np.array_equal(ref[r], order[r].cpu().numpy()[: len(ref[r])]) for r in range(16)))
    except Exception as e:
        out["cpu_oracle_error"] = str(e)[:200]
    return out


def cpu_baseline(budget_s=12.0, dev=None):
    """Oracle port of the reference's CPU path on this box's host cores, bounded sample.  Two side numbers ride
    along, both with the same port (oracle/dcnr_oracle.py = the reference's ATen op sequence): the configs[0] training
    step (B = 4096, forward + backward) on the host cores, and the ranking forward in torch EAGER on this GPU (fp32,
    TF32 off) -- what the unmodified reference would do if its device line said cuda."""
    from oracle import dcnr_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    state = make_state()
    n_req = 128
    u, i, c, x = synth_requests(n_req, CANDIDATES, 1234, "cpu")
    with torch.no_grad():
        orc.forward(state, u, i, c, x, training=False)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < budget_s:
            orc.forward(state, u, i, c, x, training=False); n += 1
    dt = time.perf_counter() - t0
    out = {"value": n * u.numel() / dt, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{n} passes over {n_req} requests x {CANDIDATES} candidates ({u.numel()} rows each), {dt:.1f} s, "
                     "torch CPU fp32, eval mode"}
    # configs[0]: training step on the host cores
    B = 4096
    g = torch.Generator().manual_seed(5)
    tu = torch.randint(0, N_USERS, (B,), generator=g); ti = torch.randint(0, N_ITEMS, (B,), generator=g)
    tc = torch.stack([torch.randint(0, k, (B,), generator=g) for k in CAT_DIMS.values()], 1)
    tx = torch.rand(B, N_NUM, generator=g); ty = (torch.rand(B, generator=g) < 0.3).float()
    orc.forward_backward(state, tu, ti, tc, tx, labels=ty, dropout_p=0.0)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 5.0:
        orc.forward_backward(state, tu, ti, tc, tx, labels=ty, dropout_p=0.0); n += 1
    dt = time.perf_counter() - t0
    out["train_step_cpu"] = {"value": n * B / dt, "unit": "samples/s", "batch": B, "ms_per_step": dt / n * 1e3,
                             "what": "configs[0]: forward + BCE + backward (dense table gradients), torch CPU fp32"}
    if dev is not None:
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            st = {k: v.to(dev) for k, v in state.items()}
            gu, gi, gc, gx = synth_requests(2048, CANDIDATES, 1234, dev)          # 1 024 000 rows per pass
            with torch.no_grad():
                for _ in range(2):
                    orc.forward(st, gu, gi, gc, gx, training=False)
                secs = time_steps(lambda: orc.forward(st, gu, gi, gc, gx, training=False), 5, 1, lambda: None)
            out["torch_eager_cuda"] = {"value": 5 * gu.numel() / secs, "unit": "candidates/s",
                                       "what": "same port in torch eager on this GPU, fp32 (allow_tf32 = False), 1 024 000 rows per pass"}
        except Exception as e:                                 # a baseline must never sink the bench line
            out["torch_eager_cuda"] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    main()
