"""CUDA-event timings of the HBM-bound kernels of the DCN-R path through the C ABI, against their algorithmic
bytes (SURVEY.md 8d / DESIGN.md 3).  python scripts/kernel_probe.py [B] -> one JSON object."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200
from dcnr_b200 import _cabi as C

P0 = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0)
N_USERS, N_ITEMS, CAT, N_NUM = 1_000_000, 100_000, {"city": 100, "hotel_type": 6}, 11


def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def probe(B, hbm_peak):
    dev = torch.device("cuda")
    m = dcnr_b200.DCN_RecSys(N_USERS, N_ITEMS, CAT, N_NUM, P0).to(dev)
    dims, ps = m._dims(), m._param_struct()
    g = torch.Generator(device=dev).manual_seed(5)
    u = torch.randint(0, N_USERS, (B,), generator=g, device=dev)
    i = torch.randint(0, N_ITEMS, (B,), generator=g, device=dev)
    c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in CAT.values()], 1).contiguous()
    x = torch.rand((B, N_NUM), generator=g, device=dev)
    batch = C.Batch(C.ptr(u), C.ptr(i), C.ptr(c), C.ptr(x), B)
    Dp, D = dims.in_dim_pad, dims.in_dim
    x0 = torch.empty((B, Dp), device=dev)
    out = {}

    def rec(name, secs, bytes_per_row, note):
        gbs = bytes_per_row * B / secs / 1e9
        out[name] = {"us": secs * 1e6, "bytes_per_row": bytes_per_row, "GB/s": gbs, "hbm_frac": gbs / hbm_peak, "note": note}

    t = timed(lambda: C.check(C.lib().dcnr_embed_concat_fwd(dims, ps, batch, C.ptr(x0), Dp, C.stream())))
    rec("K1_gather_concat_fwd", t, 32 + 44 + 128 + 228, "ids 32 + numerics 44 + user/item rows 128 + x0 write 228 (256 padded)")

    grads = [torch.empty_like(p) for p in m._ordered_params()[:4]] + [None] * (len(m._ordered_params()) - 4)
    gs = m._grad_struct(grads)
    dx = torch.randn((B, Dp), device=dev, generator=g)
    ws = torch.empty(C.lib().dcnr_workspace_bytes(dims, B, 2), dtype=torch.uint8, device=dev)
    t = timed(lambda: C.check(C.lib().dcnr_embed_scatter_bwd(dims, batch, C.ptr(dx), Dp, gs, C.ptr(ws), ws.numel(), C.stream())))
    dense_fill = (N_USERS + N_ITEMS) * 16 * 4 / B
    rec("K7_embed_scatter_bwd", t, 32 + 184 + 128, "ids 32 + dx0 read 184 + unique-row writes <= 128; EXCLUDES the dense "
        f"zero-fill of the 70 MB table gradients ({dense_fill:.0f} B/row at this batch) and the radix-sort traffic")
    out["K7_embed_scatter_bwd"]["GB/s_incl_dense_fill"] = (344 + dense_fill) * B / t / 1e9

    F_ = dcnr_b200.functional
    w = [cl.w.weight for cl in m.cross_network]; b = [cl.b for cl in m.cross_network]
    xc = torch.randn((B, D), device=dev, generator=g) * 0.1
    t = timed(lambda: F_.cross_network(xc, w, b))
    rec("K2_cross_fwd_standalone", t, 2 * D * 4, "x read + y write, unpadded rows (fused into K1 in the model: 0 extra bytes there)")
    # opt-in DCN-v2 cross layer: one tcgen05 GEMM [B,64]x[64,64] with bias / Hadamard / residual in the epilogue
    x0p = torch.randn((B, Dp), device=dev, generator=g) * 0.1
    wv2 = torch.randn((Dp, Dp), device=dev, generator=g) / 8; bv2 = torch.zeros(Dp, device=dev); yv2 = torch.empty_like(x0p)
    lws = torch.empty(max(1, C.lib().dcnr_linear_workspace_bytes(Dp, Dp, C.PRECISIONS["tf32x3"])), dtype=torch.uint8, device=dev)
    t = timed(lambda: C.check(C.lib().dcnr_cross_v2_fwd(C.ptr(x0p), Dp, C.ptr(x0p), Dp, C.ptr(wv2), Dp, C.ptr(bv2), C.ptr(yv2), Dp,
                                                       B, Dp, C.PRECISIONS["tf32x3"], C.ptr(lws), lws.numel(), C.stream())))
    rec("K2v2_cross_v2_layer_fwd", t, 3 * Dp * 4, "x0 read (Hadamard) + x read (GEMM operand; the residual box re-reads it through L2) + y write, "
        "padded 64-float rows; 2*64*64 flops per row (tf32x3)")
    del x0p, yv2
    H = 256
    z = torch.randn((B, H), device=dev, generator=g)
    mean = torch.empty(H, device=dev); rstd = torch.empty(H, device=dev)
    sc = torch.empty(C.lib().dcnr_bn_scratch_bytes(B, H), dtype=torch.uint8, device=dev)
    t = timed(lambda: C.check(C.lib().dcnr_bn_stats(C.ptr(z), H, B, H, 1e-5, 0.1, C.ptr(mean), C.ptr(rstd), None, None, None,
                                                   C.ptr(sc), sc.numel(), C.stream())))
    rec("K4_bn_stats", t, H * 4, "z read")
    ga = torch.ones(H, device=dev); be = torch.zeros(H, device=dev); o = torch.empty_like(z)
    t = timed(lambda: C.check(C.lib().dcnr_bn_act_fwd(C.ptr(z), H, C.ptr(mean), C.ptr(rstd), C.ptr(ga), C.ptr(be), None, 0, None, 0.0,
                                                     0, 0, C.ptr(o), H, B, H, C.stream())))
    rec("K4_bn_act_fwd", t, 2 * H * 4, "z read + out write")
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    res = {f"B={b}": probe(b, peaks["hbm_gbs"]) for b in (65536, B)}
    print(json.dumps(res, indent=1))
