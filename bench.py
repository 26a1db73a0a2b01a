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


def synth_module(n_users=N_USERS, n_items=N_ITEMS, params=None, seed=42):
    """A 'trained-like' DCN_RecSys (embeddings x0.1, randomised BatchNorm affine / running statistics) on the CPU;
    the product arm's own initialiser (nothing from oracle/ on this path)."""
    import dcnr_b200
    torch.manual_seed(seed)
    m = dcnr_b200.DCN_RecSys(n_users, n_items, CAT_DIMS, N_NUM, dict(P0 if params is None else params))
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
    return m


def synth_state_device(dev, seed=42):
    """The P0 model of the headline workload (1 M x 100 K tables) on the device."""
    return synth_module(seed=seed).to(dev)


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


def bind_to_gpu_numa_node(local):
    """Best effort: pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated, so the end-to-end copies do not cross the socket interconnect (8 ranks x 2.5 GB per step).  Returns the node or
    None."""
    try:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                   # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b_ = part.partition("-")
            cpus.update(range(int(a), int(b_ or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def bench_bf16_mode(model, u, i, c, x, parity_logits, rows=8192 * CANDIDATES):
    """The stated-tolerance bf16 mode of the fused tower (bf16 operands, one MMA per product) on the first `rows` candidates of
    the headline workload: throughput, and how far its logits are from the parity mode's on the same inputs.  Reported beside the
    headline, never as the headline."""
    rows = min(rows, u.numel())
    args = (u[:rows], i[:rows], c[:rows], x[:rows])
    prec, model.precision = model.precision, "bf16"
    try:
        with torch.no_grad():
            secs = time_steps(lambda: model(*args), 5, 2, lambda: None) / 5
            out = model(*args)
    finally:
        model.precision = prec
    ref = parity_logits[:rows].double()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    return {"value": rows / secs, "unit": "candidates/s", "rows": rows, "ms": secs * 1e3,
            "logits_max_abs_normalised_vs_parity_mode": err,
            "what": "DCN_RecSys(precision='bf16'): the fused tower with bf16 operands and one tcgen05 MMA per product, fp32 accumulate "
                    "and epilogue; a stated-tolerance mode (tests assert <= 3e-2 against the float64 oracle), not the default"}


def bench_request_latency(model, dev):
    """One ranking request (1 user x 500 candidates, main.py:320-325) through DCN_RecSys.eval(): device-resident inputs timed
    with CUDA events back to back, and host inputs -> host scores by wall clock (H2D of 38 KB, 2 launches, D2H of 2 KB, sync)."""
    import time
    ru, ri, rc, rx = synth_requests(1, CANDIDATES, 99, dev)
    with torch.no_grad():
        dev_us = time_steps(lambda: model(ru, ri, rc, rx), 300, 30, lambda: None) / 300 * 1e6
        hs = [t.cpu().pin_memory() for t in (ru, ri, rc, rx)]

        def host_call():
            return model(*(t.to(dev, non_blocking=True) for t in hs)).cpu()
        for _ in range(30):
            host_call()
        t0 = time.perf_counter()
        for _ in range(300):
            host_call()
        host_us = (time.perf_counter() - t0) / 300 * 1e6
    return {"rows": CANDIDATES, "us_device_resident": dev_us, "us_host_to_host": host_us,
            "what": "one request of 500 candidates through the module call; the reference's CPU forward of the same request takes "
                    "2.55 ms (SURVEY 8a, a8)"}


def load_reference():
    """The reference's own module (main.py: DCN_RecSys at main.py:93-127) from $REF_DIR or from baseline/_ref, the git-ignored
    copy __graft_entry__.build() makes in the build container and gpurun ships to the GPU box (the reference is a set of
    scripts: nothing to pip-install).  (None, None) when neither exists: the oracle port is timed instead."""
    for d in (os.environ.get("REF_DIR"), os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.exists(os.path.join(d, "main.py")):
            sys.path.insert(0, d)
            try:
                import main as ref_main                     # side effects: logging setup, set_seed(42), FastAPI app object
                return ref_main, d
            except Exception as e:                          # a missing service dependency must not sink the bench
                print(f"reference import from {d} failed: {e!r}", file=sys.stderr)
                sys.path.remove(d)
    return None, None


class ReferenceModel:
    """The unmodified reference class (kind "reference") or, when it is not importable, the oracle port of its forward /
    backward (kind "port": oracle/dcnr_oracle.py, same maths, not the same ATen op sequence)."""

    def __init__(self, state, device="cpu"):
        self.ref_main, self.src = load_reference()
        self.kind = "reference" if self.ref_main is not None else "port"
        self.device = torch.device(device)
        if self.ref_main is not None:
            self.model = self.ref_main.DCN_RecSys(N_USERS, N_ITEMS, CAT_DIMS, N_NUM, dict(P0))
            self.model.load_state_dict(state)
            self.model.to(self.device)
        else:
            self.state = {k: v.to(self.device) for k, v in state.items()}

    def infer(self, u, i, c, x):
        with torch.no_grad():
            if self.ref_main is not None:
                self.model.eval()
                return self.model(u, i, c, x)                # main.py:320-321
            from oracle import dcnr_oracle as orc
            return orc.forward(self.state, u, i, c, x, training=False)

    def train_step(self, u, i, c, x, y):
        """forward + BCEWithLogitsLoss + backward (train.py:223-225), optimizer excluded like our own 'train' number."""
        if self.ref_main is not None:
            self.model.train()
            for p in self.model.parameters():
                p.grad = None
            loss = torch.nn.BCEWithLogitsLoss()(self.model(u, i, c, x), y)
            loss.backward()
            return loss
        from oracle import dcnr_oracle as orc
        return orc.forward_backward(self.state, u, i, c, x, labels=y, dropout_p=0.0)


def reference_arm(args):
    """The reference's CPU PyTorch path -- main.DCN_RecSys itself when baseline/_ref (or $REF_DIR) holds main.py, else the
    oracle port -- on the box's host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceModel(make_state())
    n_req = 128                                                     # bounded sample per step: 64 000 rows
    u, i, c, x = synth_requests(n_req, CANDIDATES, 1234, "cpu")
    rows = u.numel()
    for _ in range(args.warmup):
        ref.infer(u, i, c, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.infer(u, i, c, x)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = f"{n_req} requests x {CANDIDATES} candidates = {rows} rows per step (of {REQUESTS * CANDIDATES})"
    _emit({
        "impl": "reference", "metric": "ranking_candidates_per_s", "value": value, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DCN-R P0 ranking inference, 1 user x 500 candidates/request (BASELINE configs[1])",
                   "tables": f"{N_USERS} users x {N_ITEMS} hotels", "sample": sample,
                   "reference_class": "main.DCN_RecSys (unmodified, eval(), torch CPU fp32)" if ref.kind == "reference"
                                      else "oracle port of main.DCN_RecSys (baseline/_ref/main.py not found)"},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": ref.kind,
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
    ap.add_argument("--precision", default=os.environ.get("DCNR_PRECISION", "fp16x3"),
                    help="dense-layer arithmetic: fp16x3 (default: fused eval tower, fp32-parity fp16 split; training runs tf32x3), "
                         "tf32x3 (per-layer tcgen05 GEMMs, fp32-parity split), fp32 (CUDA cores), tf32 / bf16 (stated tolerance)")
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
    numa_node = bind_to_gpu_numa_node(local)
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
    e2e_steps = max(2, args.steps)
    e2e_secs = max_over_ranks(time_steps(lambda: engine.score(hu, hi, hc, hx, out=hout), e2e_steps, 1, barrier))
    e2e_value = world * rows * e2e_steps / e2e_secs
    assert abs(float(hout.double().sum()) - checksum) <= 1e-6 * max(1.0, abs(checksum)) * 10, "e2e result differs"

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------
    # fp16x3 / bf16: the fused tower kernel (one launch per pass of up to 4 Mi rows: initial layer + 2R hidden layers + final dot);
    # other precisions: the per-layer dense GEMMs.  "isolated": the kernel alone on resident operands.
    M, H = min(rows, 1 << 20), P0["hidden_dim"]
    prec = C.PRECISIONS[args.precision]
    fused = args.precision in ("fp16x3", "bf16")
    if fused:
        dims, ps = model._dims(), model._param_struct()
        x0 = torch.zeros(M, dims.in_dim_pad, device=dev)
        x0[:, :dims.in_dim] = torch.randn(M, dims.in_dim, device=dev) * 0.3
        tout = torch.empty(M, device=dev); tflags = torch.zeros(4, dtype=torch.int32, device=dev)
        tws = torch.empty(C.lib().dcnr_tower_eval_workspace_bytes(dims), dtype=torch.uint8, device=dev)
        k_secs = time_steps(lambda: C.check(C.lib().dcnr_tower_eval(dims, ps, C.ptr(x0), x0.shape[1], None, C.ptr(tout), M, prec, 0,
                                                                    C.ptr(tflags), C.ptr(tws), tws.numel(), C.stream())),
                            10, 3, lambda: None)
        k_flops = M * (2.0 * dims.in_dim * H + 4 * 2.0 * H * H + 2.0 * H)
        del x0, tout
        kname = f"k_tower_eval (fused tower, tcgen05 kind::f16, {args.precision})"
        what = ("one fused-tower launch per pass of up to 4 Mi rows (initial layer + four 256x256 layers + final dot, activations in TMEM); "
                "algorithmic flops = rows x (2*57*256 + 4*2*256*256 + 2*256)")
        iso_how = f"dcnr_tower_eval on {M} resident rows alone, 10 launches after 3 warm-ups (includes the weight-pack kernel)"
    else:
        a = torch.randn(M, H, device=dev); w = torch.randn(H, H, device=dev) / 16
        sc = torch.rand(H, device=dev); sh = torch.rand(H, device=dev)
        k_secs = time_steps(lambda: dcnr_b200.functional.linear_forward_raw(a, w, sh, sc, a, True, prec), 10, 3, lambda: None)
        k_flops = 2.0 * M * H * H
        del a, w
        kname = "k_sgemm<true,true> (fp32 CUDA-core)" if args.precision == "fp32" else f"k_gemm_tc, tcgen05 ({args.precision})"
        what = ("per 1 Mi-row chunk one 256x64 initial layer + four 256x256 layers with the folded-BN / residual / ReLU epilogues; "
                "algorithmic flops = 2 m n k with the initial layer counted at its unpadded width")
        iso_how = f"dcnr_linear_fwd {M}x{H}x{H} + scale/shift/residual/relu alone, 10 launches after 3 warm-ups"
    isolated = k_flops * 10 / k_secs / 1e12
    # live: a repeat of the timed steps with a CUDA event pair (on the launching stream) around EVERY dense-layer kernel launch
    C.gemm_timing_begin()
    t_rep = time_steps(step_resident, args.steps, 0, lambda: None)
    g_ms, g_n, g_fl = C.gemm_timing_end()
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    terms = 3 if args.precision in ("fp16x3", "tf32x3") else 1
    pipe_peak = pk["tensor"] / (2 if args.precision in ("tf32x3", "tf32") else 1)      # kind::tf32 runs at half the kind::f16 rate
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": achieved / pk["tensor"], "traffic": None, "kernel": kname,
                "launches": g_n, "avg_launch_us": g_ms * 1e3 / max(g_n, 1), "share_of_step": g_ms * 1e-3 / t_rep,
                "mma_per_algorithmic_mac": terms,
                "frac_vs_split_ceiling": achieved / (pipe_peak / terms) if args.precision != "fp32" else None,
                "split_ceiling_tflops": pipe_peak / terms if args.precision != "fp32" else None,
                "how": f"CUDA event pairs on the launching stream around every dense-layer kernel launch over a repeat of the "
                       f"{args.steps} timed steps ({g_n} launches: {what}); achieved = sum of algorithmic flops / sum of launch "
                       f"durations; peak = bf16 sustained ({pk['src']}); frac_vs_split_ceiling = achieved / (tensor-pipe peak of the "
                       f"MMA kind / MMAs per algorithmic MAC): the {terms}-term parity split issues {terms} MMAs per product, so its "
                       f"own ceiling is peak / {terms}" + (" of the kind::tf32 rate (half the measured bf16 rate)" if "tf32" in args.precision else ""),
                "isolated_tflops": isolated, "isolated_how": iso_how,
                "step_algorithmic_tflops": FLOP_PER_ROW * rows * args.steps / secs / 1e12}

    result = {
        "metric": "ranking_candidates_per_s", "value": value, "unit": "candidates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-parity split, fp32 accumulate)", "tf32": "tf32", "bf16": "bf16",
                  "fp16x3": "fp16x3 (fp32-parity 3-term fp16 split on kind::f16 MMAs, fp32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "DCN-R P0 ranking inference, 1 user x 500 candidates/request, "
                               f"{args.requests} requests/step per GPU (BASELINE configs[1])",
                   "rows_per_step_per_gpu": rows, "tables": f"{N_USERS} users x {N_ITEMS} hotels",
                   "model": "emb16 hidden256 cross3 res2 (P0)", "l2": "inputs 2.5 GB per step > 126 MB L2, no flush needed",
                   "parallelism": f"requests sharded x{world}, no data-path collective",
                   "weights": "the fused tower's fp16 hi/lo weight pack (1.3 MB) is rebuilt only when a parameter changes "
                              "(k_tower_prep, 38 us = 0.07 % of a step); every step runs the gather and the whole tower on fresh inputs",
                   "host_numa_node_of_rank0": numa_node},
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
        result["request_latency"] = bench_request_latency(model, dev)
        if world == 1 and args.precision == "fp16x3":
            result["bf16_mode"] = bench_bf16_mode(model, u, i, c, x, out_dev["logits"])
        comm = None
        if world > 1:
            comm = dcnr_b200.distributed.Communicator()
        result["train"] = bench_train(model, dev, world, rank, comm, barrier, max_over_ranks, single_ref=True)
        extras = {}
        if world > 1:      # the same data-parallel step with 65 536 rows PER GPU: what the sync costs when every GPU has real work
            result["train"]["parity_vs_single_device"] = train_parity_vs_single_device(model, dev, world, rank, comm, barrier, 65_536)
            tw = bench_train(model, dev, world, rank, comm, barrier, max_over_ranks, global_batch=65_536 * world, steps=5)
            tw["scaling"] = "weak"
            result["train_weak"] = tw
            if rank == 0 and "single_device" in result["train"]:
                t1 = result["train"]["single_device"]["ms_per_step"]
                extras["train_strong_eff"] = t1 / result["train"]["ms_per_step"] / world
                extras["train_weak_eff"] = t1 / tw["ms_per_step"]
        del model
        torch.cuda.empty_cache()
        result["sharded"] = bench_sharded(dev, world, rank, comm, barrier, max_over_ranks, args.precision)
        if world > 1 and rank == 0 and "single_device" in result["sharded"]:
            extras["sharded_eff"] = result["sharded"]["single_device"]["ms_per_step"] / result["sharded"]["ms_per_step"] / world
        if extras:
            extras["how"] = ("efficiency = (single-device time of the same global batch, timed on rank 0 inside this run) / "
                             "(N x N-rank time) for the strong-scaling steps; weak: single-device 65 536-row step / N-rank step "
                             "with 65 536 rows per GPU")
            result["scaling_extras"] = extras
        torch.cuda.empty_cache()
        result["similarity"] = bench_similarity(dev, pk, world, rank, barrier, max_over_ranks)
        torch.cuda.empty_cache()
        if world == 1:
            result["mmr"] = bench_mmr(dev)
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


def global_train_batch(global_batch, dev, seed=99):
    """ONE seeded global batch, generated identically on every rank (each rank then takes its contiguous slice), so an
    N-rank step can be compared with the single-device step on the same rows."""
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.randint(0, N_USERS, (global_batch,), generator=g, device=dev)
    i = torch.randint(0, N_ITEMS, (global_batch,), generator=g, device=dev)
    c = torch.stack([torch.randint(0, n, (global_batch,), generator=g, device=dev) for n in CAT_DIMS.values()], 1)
    x = torch.rand((global_batch, N_NUM), generator=g, device=dev)
    y = (torch.rand(global_batch, generator=g, device=dev) < 0.3).float()
    return u, i, c, x, y


def _norm_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def train_parity_vs_single_device(model, dev, world, rank, comm, barrier, global_batch):
    """The N-rank data-parallel step against the SAME step run on rank 0 alone over the concatenated batch (dropout off so
    both draw no mask): logits of this rank's slice and every gradient tensor after the exchange, max-abs-normalised.
    Returns the worst error and where it sits (rank 0's view; the gradients are identical on every rank by construction)."""
    import dcnr_b200
    from dcnr_b200.distributed import allreduce_gradients, attach, shard_range
    u, i, c, x, y = global_train_batch(global_batch, dev)
    b0, b1 = shard_range(global_batch, rank, world)
    p_drop, model._shape["dropout"] = model._shape["dropout"], 0.0
    bufs = {n: b.clone() for n, b in model.named_buffers()}
    model.train()
    out = {}
    try:
        def step(sl, use_comm):
            attach(model, comm if use_comm else None)
            for p in model.parameters():
                p.grad = None
            logits = model(u[sl], i[sl], c[sl], x[sl])
            _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y[sl])
            logits.backward(gradient=dl * ((b1 - b0) / global_batch) if use_comm else dl)     # mean over the GLOBAL batch
            if use_comm:
                allreduce_gradients(model.parameters_to_allreduce(), comm=comm, average=False)
            return logits.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}
        lo_dp, g_dp = step(slice(b0, b1), True)
        model.load_state_dict({**model.state_dict(), **bufs})           # the step moved the running statistics
        barrier()
        if rank == 0:
            lo_1, g_1 = step(slice(0, global_batch), False)
            scale = max(float(g.abs().max()) for g in g_1.values())
            errs = {"logits": _norm_err(lo_dp, lo_1[b0:b1])}
            for n in g_1:
                if ".layer1.bias" in n or ".layer2.bias" in n:          # cancelled by the following BatchNorm: true gradient is 0
                    errs[n] = float((g_dp[n] - g_1[n]).abs().max()) / scale
                else:
                    errs[n] = _norm_err(g_dp[n], g_1[n])
            worst = max(errs, key=errs.get)
            out = {"max_err": errs[worst], "worst_tensor": worst, "logits_err": errs["logits"], "tensors_compared": len(errs),
                   "what": f"{world}-rank step (global-batch BatchNorm, sparse table-gradient exchange, all-reduce) vs the same "
                           f"{global_batch}-row step on rank 0 alone, dropout off, max-abs-normalised"}
        barrier()
    finally:
        model._shape["dropout"] = p_drop
        model.load_state_dict({**model.state_dict(), **bufs})
        attach(model, None)
        model.eval()
    return out


def bench_train(model, dev, world, rank, comm, barrier, max_over_ranks, global_batch=65_536, steps=10, single_ref=False):
    """BASELINE configs[2]: P0 training step (fwd + BCE + bwd + gradient all-reduce), global batch 65 536
    split over the ranks (strong scaling), 1 M x 100 K tables, dropout 0.6, dense embedding grads.  With N > 1
    the BatchNorm statistics cover the global batch (library communicator), the loss gradient is scaled by the
    global batch and the gradients are SUMMED over ranks: the step is the single-device step on 65 536 rows.
    Every rank slices ONE seeded global batch.  single_ref (N > 1): rank 0 also times the whole global batch alone,
    which gives the scaling efficiency of this very run."""
    import dcnr_b200
    from dcnr_b200 import _cabi as C
    from dcnr_b200.distributed import allreduce_gradients, attach, shard_range
    gu, gi, gc, gx, gy = global_train_batch(global_batch, dev)
    b0, b1 = shard_range(global_batch, rank, world)
    B = b1 - b0
    u, i, c, x, y = (t[b0:b1].contiguous() for t in (gu, gi, gc, gx, gy))
    model.train()
    attach(model, comm)
    params = list(model.parameters())

    def step():
        for p in params:
            p.grad = None
        logits = model(u, i, c, x)
        _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
        if world > 1:
            dl = dl * (B / global_batch)                      # mean over the GLOBAL batch
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
    single = None
    if single_ref and world > 1:
        attach(model, None)
        if rank == 0:
            def step1():
                for p in params:
                    p.grad = None
                logits = model(gu, gi, gc, gx)
                _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), gy)
                logits.backward(gradient=dl)
            s1 = time_steps(step1, steps, 3, lambda: None)
            single = {"ms_per_step": s1 / steps * 1e3, "value": global_batch * steps / s1, "unit": "samples/s",
                      "what": f"the same {global_batch}-row step on rank 0 alone, timed inside this run"}
        barrier()
    model.eval()
    attach(model, None)
    flops = 3 * FLOP_PER_ROW * global_batch
    out = {"metric": "train_samples_per_s", "value": global_batch * steps / secs, "unit": "samples/s",
           "global_batch": global_batch, "per_gpu_batch": B, "ms_per_step": secs / steps * 1e3, "scaling": "strong",
           "algorithmic_tflops": flops * steps / secs / 1e12, "cuda_graph": graphed,
           "includes": "forward + BCE + backward + NCCL all-reduce of the dense gradients; when N > 1 also global-batch "
                       "BatchNorm and the all-gather of (id, gradient row) pairs that builds the table gradients "
                       "(optimizer excluded)",
           "gpu_launches_per_step": per_step}
    if comm is not None and world > 1:
        out["syncbn_exchange"] = ("one peer-to-peer kernel per BatchNorm exchange (CUDA IPC over NVLink)" if comm.uses_peer_memory
                                  else "NCCL all-gather per BatchNorm exchange")
    if single is not None:
        out["single_device"] = single
    return out


def sharded_parity_vs_single_device(dev, world, rank, comm, barrier, precision, rows=400_000, per_rank=8_192):
    """RowShardedDCN over `world` ranks against DCN_RecSys with the full tables on rank 0, same seeded global batch, dropout
    off: logits of rank 0's slice, its table shards' gradients (rows rank 0 owns) and the dense gradients."""
    import dcnr_b200
    from dcnr_b200.distributed import RowShardedDCN, allreduce_gradients, attach, shard_range
    params = dict(P0); params["dropout"] = 0.0
    state = synth_module(rows, rows // 4, params, seed=3).state_dict()
    GB = per_rank * world
    g = torch.Generator().manual_seed(17)
    u = torch.randint(0, rows, (GB,), generator=g).to(dev); i = torch.randint(0, rows // 4, (GB,), generator=g).to(dev)
    c = torch.stack([torch.randint(0, n, (GB,), generator=g) for n in CAT_DIMS.values()], 1).to(dev)
    x = torch.rand(GB, N_NUM, generator=g).to(dev); y = (torch.rand(GB, generator=g) < 0.3).float().to(dev)
    b0, b1 = shard_range(GB, rank, world)
    sh = RowShardedDCN(rows, rows // 4, CAT_DIMS, N_NUM, params, comm, precision=precision, device=dev)
    core = {k: v for k, v in state.items() if not k.startswith(("user_embedding", "item_embedding"))}
    sh.core.load_state_dict({**core, "user_embedding.weight": torch.zeros(1, 16), "item_embedding.weight": torch.zeros(1, 16)})
    sh.core.to(dev)
    sh.user_table.load_full(state["user_embedding.weight"]); sh.item_table.load_full(state["item_embedding.weight"])
    sh.train()
    lo = sh(u[b0:b1], i[b0:b1], c[b0:b1], x[b0:b1])
    _, dl = dcnr_b200.functional.bce_with_logits(lo.detach(), y[b0:b1])
    lo.backward(gradient=dl * ((b1 - b0) / GB))
    allreduce_gradients(sh.dense_parameters(), comm=comm, average=False)
    out = {}
    barrier()
    if rank == 0:
        full = dcnr_b200.DCN_RecSys(rows, rows // 4, CAT_DIMS, N_NUM, params, precision=precision)
        full.load_state_dict(state)
        full.to(dev).train()
        lf = full(u, i, c, x)
        _, dlf = dcnr_b200.functional.bce_with_logits(lf.detach(), y)
        lf.backward(gradient=dlf)
        errs = {"logits": _norm_err(lo.detach(), lf.detach()[b0:b1]),
                "user_shard": _norm_err(sh.user_table.weight.grad, full.user_embedding.weight.grad[rank::world]),
                "item_shard": _norm_err(sh.item_table.weight.grad, full.item_embedding.weight.grad[rank::world])}
        skip = ("user_embedding", "item_embedding")
        scale = max(float(p.grad.abs().max()) for n, p in full.named_parameters() if not n.startswith(skip))
        for (n, pf), (_, ps) in zip(((n, p) for n, p in full.named_parameters() if not n.startswith(skip)),
                                    ((n, p) for n, p in sh.core.named_parameters() if not n.startswith(skip))):
            if ".layer1.bias" in n or ".layer2.bias" in n:
                errs[n] = float((ps.grad - pf.grad).abs().max()) / scale
            else:
                errs[n] = _norm_err(ps.grad, pf.grad)
        worst = max(errs, key=errs.get)
        out = {"max_err": errs[worst], "worst_tensor": worst, "logits_err": errs["logits"], "tensors_compared": len(errs),
               "what": f"row-sharded step over {world} ranks ({rows}-row tables, {GB} rows) vs the unsharded model on rank 0, "
                       "dropout off, max-abs-normalised"}
        del full
    barrier()
    del sh
    return out


def bench_sharded(dev, world, rank, comm, barrier, max_over_ranks, precision, global_batch=262_144, rows=100_000_000,
                  steps=5):
    """BASELINE configs[4]: user / item tables of 100 M rows x 16 (6.4 GB each) row-sharded over the ranks by
    row % N, global batch 262 144 (ONE seeded batch, sliced by rank), forward + backward with the NCCL all-to-all
    exchange of ids / rows / gradient rows and the owner-side sorted-segment scatter-add; dense part data parallel.
    N > 1: rank 0 also times the whole configuration alone (tables unsharded on one GPU) for the scaling efficiency,
    and a smaller instance is compared with the unsharded model (parity_vs_single_device)."""
    import dcnr_b200
    from dcnr_b200.distributed import Communicator, RowShardedDCN, allreduce_gradients, shard_range
    own = comm is None
    if own:
        comm = Communicator()                                 # world 1: no NCCL traffic, same code path
    g = torch.Generator(device=dev).manual_seed(321)
    gu = torch.randint(0, rows, (global_batch,), generator=g, device=dev)
    gi = torch.randint(0, rows, (global_batch,), generator=g, device=dev)
    gc = torch.stack([torch.randint(0, n, (global_batch,), generator=g, device=dev) for n in CAT_DIMS.values()], 1)
    gx = torch.rand((global_batch, N_NUM), generator=g, device=dev)
    gy = (torch.rand(global_batch, generator=g, device=dev) < 0.3).float()

    def run(cm, w, sl):
        model = RowShardedDCN(rows, rows, CAT_DIMS, N_NUM, P0, cm, precision=precision, device=dev).train()
        with torch.no_grad():
            model.user_table.weight.mul_(0.1); model.item_table.weight.mul_(0.1)
        u, i, c, x, y = (t[sl].contiguous() for t in (gu, gi, gc, gx, gy))
        dense = model.dense_parameters()
        every = list(model.parameters())

        def step():
            for p in every:
                p.grad = None
            logits = model(u, i, c, x)
            _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
            logits.backward(gradient=dl * (u.numel() / global_batch))
            allreduce_gradients(dense, comm=cm, average=False)
        if w == world:
            secs = max_over_ranks(time_steps(step, steps, 2, barrier))
        else:
            secs = time_steps(step, steps, 2, lambda: None)
        exch = model.user_table.last_exchange_bytes + model.item_table.last_exchange_bytes
        rows_here = model.user_table.weight.shape[0]
        del model
        torch.cuda.empty_cache()
        return secs, exch, rows_here
    b0, b1 = shard_range(global_batch, rank, world)
    secs, exch, rows_here = run(comm, world, slice(b0, b1))
    out = {"metric": "sharded_train_samples_per_s", "value": global_batch * steps / secs, "unit": "samples/s",
           "global_batch": global_batch, "per_gpu_batch": b1 - b0, "table_rows": rows, "rows_per_rank": rows_here,
           "ms_per_step": secs / steps * 1e3, "scaling": "strong",
           "alltoall_bytes_per_rank_fwd": exch, "includes": "ids/rows all-to-all fwd, gradient rows all-to-all + owner-side "
           "scatter-add bwd (dense shard gradients, zero-filled every step), dense all-reduce; optimizer excluded"}
    if world > 1:
        import torch.distributed as dist
        solo_group = dist.new_group([0])                      # collective: every rank calls it
        if rank == 0:
            solo = Communicator(group=solo_group)
            s1, _, _ = run(solo, 1, slice(0, global_batch))
            solo.close()
            out["single_device"] = {"ms_per_step": s1 / steps * 1e3, "value": global_batch * steps / s1, "unit": "samples/s",
                                    "what": "the same configuration with both 100 M-row tables on rank 0 alone, timed inside this run"}
        barrier()
        out["parity_vs_single_device"] = sharded_parity_vs_single_device(dev, world, rank, comm, barrier, precision)
    if own:
        comm.close()
    return out


def bench_similarity(dev, pk, world=1, rank=0, barrier=lambda: None, max_over_ranks=lambda v: v, n=10_000_000, d=16, k=201):
    """BASELINE configs[3]: cosine top-201 over a 10 M x 16 catalog, query batches Q in {1, 8, 32, 1024}.
    One GPU: 640 MB scan per catalog pass; hbm_frac = one catalog read / whole-call time (normalise + scan + merge).  N > 1: the catalog is split into contiguous shards (10 M / N rows per GPU), every rank
    scans its shard, the (dist, idx) lists are all-gathered and merged in the contract order
    (distributed.ShardedNearestNeighbors); times are max over ranks, and rank 0 checks the Q = 32 answer bit for bit against
    the unsharded catalog on one GPU."""
    import dcnr_b200
    from dcnr_b200.distributed import ShardedNearestNeighbors, shard_range
    g = torch.Generator(device=dev).manual_seed(7)
    E = torch.randn(n, d, device=dev, generator=g)             # same seed on every rank: the same global catalog
    qidx = torch.randint(0, n, (1024,), device=dev, generator=g)
    out = {}
    if world == 1:
        model = dcnr_b200.NearestNeighbors().fit(E)
        from dcnr_b200 import _cabi as C
        for nq in (1, 8, 32, 1024):
            Q = E[qidx[:nq]].contiguous()
            reps = 10 if nq <= 32 else 3
            tc = nq >= model.tc_min_queries and bool(C.lib().dcnr_knn_tc_supported(n, d, nq, k))
            secs = time_steps(lambda: model.kneighbors_tensor(Q, k), reps, 3, lambda: None) / reps
            passes = 1 if (nq == 1 or tc) else (nq + 7) // 8
            out[f"q{nq}"] = {"ms": secs * 1e3, "pairs_per_s": nq * n / secs, "catalog_passes": passes,
                             "path": "tensor-core shortlist + exact re-score (dcnr_knn_topk_tc)" if tc else "exact streaming scan (dcnr_knn_topk)",
                             "hbm_frac": (n * d * 4) / secs / 1e9 / pk["hbm"]}
        out["how"] = ("whole kneighbors_tensor call (normalise + scan + select, incl. the status read-back of the tensor-core "
                      "path), CUDA events over back-to-back calls")
        return {"metric": "similarity_query_candidate_pairs_per_s", "catalog": f"{n} x {d} f32", "k": k, **out}
    s0, s1 = shard_range(n, rank, world)
    snn = ShardedNearestNeighbors(n_neighbors=k).fit_shard(E[s0:s1].contiguous(), s0)
    exact = None
    for nq in (1, 32, 1024):
        Q = E[qidx[:nq]].contiguous()
        reps = 10 if nq <= 32 else 3
        secs = max_over_ranks(time_steps(lambda: snn.kneighbors_tensor(Q, k), reps, 2, barrier)) / reps
        out[f"q{nq}"] = {"ms": secs * 1e3, "pairs_per_s": nq * n / secs, "rows_per_gpu": s1 - s0,
                         "allgather_bytes_per_rank": nq * k * 12}
        if nq == 32:
            ds, is_ = snn.kneighbors_tensor(Q, k)
            barrier()
            if rank == 0:
                df, if_ = dcnr_b200.NearestNeighbors().fit(E).kneighbors_tensor(Q, k)
                exact = bool(torch.equal(is_, if_)) and bool(torch.equal(ds, df))
            barrier()
    return {"metric": "similarity_query_candidate_pairs_per_s", "catalog": f"{n} x {d} f32 in {world} contiguous shards", "k": k,
            "sharded": True, "bit_exact_vs_unsharded_q32": exact,
            "includes": "per-shard normalise + scan + top-k, torch.distributed all_gather of (dist, idx), cross-shard merge", **out}


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
    """The reference's CPU path on this box's host cores, bounded sample: main.DCN_RecSys itself (kind "reference") when
    baseline/_ref holds main.py, else the oracle port.  Side numbers with the same class: the configs[0] training step
    (B = 4096, forward + BCE + backward) on the host cores, and -- what the unmodified reference would do if its device line
    said cuda -- the ranking forward and the configs[2] training step in torch EAGER on this GPU
    (fp32, TF32 off)."""
    torch.set_num_threads(os.cpu_count() or 1)
    state = make_state()
    ref = ReferenceModel(state)
    n_req = 128
    u, i, c, x = synth_requests(n_req, CANDIDATES, 1234, "cpu")
    ref.infer(u, i, c, x)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < budget_s:
        ref.infer(u, i, c, x); n += 1
    dt = time.perf_counter() - t0
    out = {"value": n * u.numel() / dt, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": ref.kind,
           "sample": f"{n} passes over {n_req} requests x {CANDIDATES} candidates ({u.numel()} rows each), {dt:.1f} s, "
                     "torch CPU fp32, eval mode"}
    # configs[0]: training step on the host cores
    B = 4096
    g = torch.Generator().manual_seed(5)
    tu = torch.randint(0, N_USERS, (B,), generator=g); ti = torch.randint(0, N_ITEMS, (B,), generator=g)
    tc = torch.stack([torch.randint(0, k, (B,), generator=g) for k in CAT_DIMS.values()], 1)
    tx = torch.rand(B, N_NUM, generator=g); ty = (torch.rand(B, generator=g) < 0.3).float()
    ref.train_step(tu, ti, tc, tx, ty)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 5.0:
        ref.train_step(tu, ti, tc, tx, ty); n += 1
    dt = time.perf_counter() - t0
    out["train_step_cpu"] = {"value": n * B / dt, "unit": "samples/s", "batch": B, "ms_per_step": dt / n * 1e3, "kind": ref.kind,
                             "what": "configs[0]: forward + BCE + backward (dense table gradients), torch CPU fp32"}
    if dev is not None:
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            gref = ReferenceModel(state, dev)
            gu, gi, gc, gx = synth_requests(2048, CANDIDATES, 1234, dev)          # 1 024 000 rows per pass
            for _ in range(2):
                gref.infer(gu, gi, gc, gx)
            secs = time_steps(lambda: gref.infer(gu, gi, gc, gx), 5, 1, lambda: None)
            out["torch_eager_cuda"] = {"value": 5 * gu.numel() / secs, "unit": "candidates/s", "kind": gref.kind,
                                       "what": "the reference class in torch eager on this GPU, fp32 (allow_tf32 = False), 1 024 000 rows per pass"}
            B2 = 65_536
            g2 = torch.Generator(device=dev).manual_seed(99)
            tu = torch.randint(0, N_USERS, (B2,), generator=g2, device=dev); ti = torch.randint(0, N_ITEMS, (B2,), generator=g2, device=dev)
            tc = torch.stack([torch.randint(0, k, (B2,), generator=g2, device=dev) for k in CAT_DIMS.values()], 1)
            tx = torch.rand((B2, N_NUM), generator=g2, device=dev); ty = (torch.rand(B2, generator=g2, device=dev) < 0.3).float()
            for _ in range(2):
                gref.train_step(tu, ti, tc, tx, ty)
            secs = time_steps(lambda: gref.train_step(tu, ti, tc, tx, ty), 10, 1, lambda: None)
            out["torch_eager_cuda_train"] = {"value": 10 * B2 / secs, "unit": "samples/s", "batch": B2, "ms_per_step": secs / 10 * 1e3,
                                             "kind": gref.kind,
                                             "what": "configs[2] shape: the reference class, forward + BCE + backward in torch eager on this "
                                                     "GPU (fp32, allow_tf32 = False, dropout 0.6, dense table gradients, optimizer excluded)"}
        except Exception as e:                                 # a baseline must never sink the bench line
            out["torch_eager_cuda_error"] = str(e)[:200]
    return out


if __name__ == "__main__":
    main()
