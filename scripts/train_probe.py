"""One P0 training step (fwd + BCE + bwd) at B=65536 on one GPU, timed with CUDA events: python scripts/train_probe.py [precision] [steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200, bench
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
model = bench.synth_state_device(dev)
model.precision = prec
model.train()
B = 65536
g = torch.Generator(device=dev).manual_seed(99)
u = torch.randint(0, bench.N_USERS, (B,), generator=g, device=dev)
i = torch.randint(0, bench.N_ITEMS, (B,), generator=g, device=dev)
c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in bench.CAT_DIMS.values()], 1)
x = torch.rand((B, bench.N_NUM), generator=g, device=dev)
y = (torch.rand(B, generator=g, device=dev) < 0.3).float()
params = list(model.parameters())
def step():
    for p in params: p.grad = None
    logits = model(u, i, c, x)
    _, dl = dcnr_b200.functional.bce_with_logits(logits.detach(), y)
    logits.backward(gradient=dl)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train step {prec} B={B}: {ms:.3f} ms  {B/ms/1e3:.2f} M samples/s  {1.662e6*B/ms/1e9:.1f} TFLOP/s algorithmic")
