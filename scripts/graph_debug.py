"""Eager vs graphed training step: per-tensor gradient differences on one batch (debug aid)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dcnr_b200
from oracle import dcnr_oracle as orc
from tests.helpers import synth_inputs
n_users, n_items, cat_dims, n_num = 3000, 900, {"city": 100, "hotel_type": 6}, 11
params = dict(emb_dim=16, hidden_dim=256, n_cross_layers=3, n_res_blocks=2, dropout=0.0)
state = orc.make_state(n_users, n_items, cat_dims, n_num, params, seed=4, emb_scale=0.1, randomize_bn=True)
B = 2048
batches = [tuple(t.cuda() for t in synth_inputs(n_users, n_items, cat_dims, n_num, B, seed=50 + s)) for s in range(3)]
def fresh():
    m = dcnr_b200.DCN_RecSys(n_users, n_items, cat_dims, n_num, params, precision="tf32x3"); m.load_state_dict(state); return m.cuda().train()
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
e1, e2, gm = fresh(), fresh(), fresh()
gs = dcnr_b200.training.GraphedTrainStep(gm, B); gs.load(*batches[0]); gs.capture()
for s, (u, i, c, x, y) in enumerate(batches):
    for p in e1.parameters(): p.grad = None
    torch.nn.BCEWithLogitsLoss()(e1(u, i, c, x), y).backward()
    for p in e2.parameters(): p.grad = None
    lo = e2(u, i, c, x); _, dl = dcnr_b200.functional.bce_with_logits(lo.detach(), y); lo.backward(gradient=dl)
    gs(u, i, c, x, y)
    worst = {}
    for (n, a), (_, b), (_, g) in zip(e1.named_parameters(), e2.named_parameters(), gm.named_parameters()):
        worst[n] = (rel(a.grad, b.grad), rel(g.grad, b.grad))
    w1 = max(worst, key=lambda k: worst[k][0]); w2 = max(worst, key=lambda k: worst[k][1])
    print(f"step {s}: torch-BCE eager vs fused-BCE eager: worst {w1} {worst[w1][0]:.2e};  graph vs fused-BCE eager: worst {w2} {worst[w2][1]:.2e}")
for (n, b1), (_, b2) in zip(e2.named_buffers(), gm.named_buffers()):
    if b1.dtype.is_floating_point: print(n, rel(b2, b1))
    else: print(n, int(b1), int(b2))
