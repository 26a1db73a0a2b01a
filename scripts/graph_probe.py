"""GraphedTrainStep vs the eager step on one GPU: same gradients (dropout 0), timing of both."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcnr_b200, bench
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = bench.synth_state_device(dev); model.precision = "tf32x3"
for blk in model.res_blocks: blk.dropout.p = 0.0
model._shape["dropout"] = 0.0
g = torch.Generator(device=dev).manual_seed(99)
u = torch.randint(0, bench.N_USERS, (B,), generator=g, device=dev); i = torch.randint(0, bench.N_ITEMS, (B,), generator=g, device=dev)
c = torch.stack([torch.randint(0, n, (B,), generator=g, device=dev) for n in bench.CAT_DIMS.values()], 1)
x = torch.rand((B, bench.N_NUM), generator=g, device=dev); y = (torch.rand(B, generator=g, device=dev) < 0.3).float()
state = {k: v.clone() for k, v in model.state_dict().items()}
model.train()
params = list(model.parameters())
def eager():
    for p in params: p.grad = None
    lo = model(u, i, c, x); _, dl = dcnr_b200.functional.bce_with_logits(lo.detach(), y); lo.backward(gradient=dl)
eager(); ref = [p.grad.clone() for p in params]
for _ in range(3): eager()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [eager() for _ in range(20)]; e1.record(); torch.cuda.synchronize(); t_eager = e0.elapsed_time(e1) / 20
model.load_state_dict(state)
gs = dcnr_b200.training.GraphedTrainStep(model, B); gs.load(u, i, c, x, y); gs.capture()
model.load_state_dict(state)                      # warm-up / capture moved the running statistics
gs()
err = max(float((p.grad - r).abs().max() / r.abs().max().clamp_min(1e-30)) for p, r in zip(params, ref) if r.abs().max() > 1e-10)
torch.cuda.synchronize(); e0.record(); [gs() for _ in range(50)]; e1.record(); torch.cuda.synchronize(); t_graph = e0.elapsed_time(e1) / 50
print(f"B={B}: eager {t_eager:.3f} ms, graph {t_graph:.3f} ms, max grad diff graph vs eager {err:.2e}, loss {float(gs.loss):.5f}")
