"""Times / profiles one dense layer through the C ABI: python scripts/gemm_probe.py <precision> [M] [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
H = 256
a = torch.randn(M, H, device="cuda"); w = torch.randn(H, H, device="cuda") / 16
sc = torch.rand(H, device="cuda"); sh = torch.rand(H, device="cuda")
F_ = dcnr_b200.functional
for variant, fn in (("full epilogue", lambda: F_.linear_forward_raw(a, w, sh, sc, a, True, prec)),
                    ("bias only", lambda: F_.linear_forward_raw(a, w, sh, None, None, False, prec))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{prec} {variant}: {ms:.3f} ms  {2*M*H*H/ms/1e9:.1f} TFLOP/s  {(3 if 'full' in variant else 2)*M*H*4/ms/1e6:.0f} GB/s")
