"""K7 (embedding-gradient scatter) alone through the C ABI: python scripts/k7_probe.py [B] [reps]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200
from dcnr_b200 import _cabi as C
from kernel_probe import P0, N_USERS, N_ITEMS, CAT, N_NUM, timed

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda")
m = dcnr_b200.DCN_RecSys(N_USERS, N_ITEMS, CAT, N_NUM, P0).to(dev)
dims, ps = m._dims(), m._param_struct()
g = torch.Generator(device=dev).manual_seed(5)
u = torch.randint(0, N_USERS, (B,), generator=g, device=dev)
i = torch.randint(0, N_ITEMS, (B,), generator=g, device=dev)
c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in CAT.values()], 1).contiguous()
x = torch.rand((B, N_NUM), generator=g, device=dev)
batch = C.Batch(C.ptr(u), C.ptr(i), C.ptr(c), C.ptr(x), B)
Dp = dims.in_dim_pad
grads = [torch.empty_like(p) for p in m._ordered_params()[:4]] + [None] * (len(m._ordered_params()) - 4)
gs = m._grad_struct(grads)
dx = torch.randn((B, Dp), device=dev, generator=g)
ws = torch.empty(C.lib().dcnr_workspace_bytes(dims, B, 2), dtype=torch.uint8, device=dev)
t = timed(lambda: C.check(C.lib().dcnr_embed_scatter_bwd(dims, batch, C.ptr(dx), Dp, gs, C.ptr(ws), ws.numel(), C.stream())),
          reps=reps, warm=max(2, reps))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
fill = (N_USERS + N_ITEMS) * 16 * 4
print(f"K7 B={B}: {t*1e6:.1f} us  {(344*B+fill)/t/1e9:.0f} GB/s algorithmic (344 B/row + {fill/1e6:.0f} MB dense fill)  {(344*B+fill)/t/1e9/peak:.3f} of {peak:.0f} GB/s")
