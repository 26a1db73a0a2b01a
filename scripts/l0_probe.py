import sys, os, torch
sys.path.insert(0, "/root/repo")
import dcnr_b200
F_ = dcnr_b200.functional
M = 1 << 20
a = torch.randn(M, 64, device="cuda"); w = torch.randn(256, 64, device="cuda") / 8; b = torch.rand(256, device="cuda")
for _ in range(2): F_.linear_forward_raw(a, w, b, None, None, False, "tf32x3")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): F_.linear_forward_raw(a, w, b, None, None, False, "tf32x3")
e1.record(); torch.cuda.synchronize()
print("L0-shape 1Mx256x64 tf32x3: %.3f ms" % (e0.elapsed_time(e1) / 10))
