"""K1 (gather + concat) alone through the C ABI: python scripts/k1_probe.py [B] [reps] -> one line (us, GB/s, fraction of the HBM peak)"""
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
x0 = torch.empty((B, Dp), device=dev)
t = timed(lambda: C.check(C.lib().dcnr_embed_concat_fwd(dims, ps, batch, C.ptr(x0), Dp, C.stream())), reps=reps, warm=max(2, reps))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
print(f"K1 B={B}: {t*1e6:.1f} us  {432*B/t/1e9:.0f} GB/s algorithmic (432 B/row)  {432*B/t/1e9/peak:.3f} of {peak:.0f} GB/s")
