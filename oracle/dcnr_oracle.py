"""CPU oracle for the DCN-R hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's DCN-R model so
that the CUDA path can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product package never does (it has no CPU fallback).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the reference's own
``main.DCN_RecSys`` (imported from /root/reference in the build container) and
stores its inputs, state_dict, logits and gradients under ``tests/golden/``;
``tests/test_oracle.py`` checks every function below against those vectors.

Two restatements live here:

* ``forward`` / ``forward_backward`` -- a *functional* interpreter of a
  reference ``state_dict`` built from torch CPU ops (the same ATen kernels the
  reference dispatches), any float dtype.  Used as the parity checker and as the
  "port" CPU baseline in bench.py.
* ``np_forward_backward`` -- closed-form numpy/float64 forward + hand-derived
  backward.  This is the algorithm the CUDA kernels implement line by line
  (cross-layer closed form, BatchNorm batch statistics, sorted-segment embedding
  scatter), so it documents the kernels' maths and is itself checked against the
  reference's autograd in the golden tests.

Reference lines followed (all in /root/reference):
  embedding widths            train.py:136-141   main.py:102-107
  x0 column order             train.py:156-159   main.py:116-119
  CrossLayer                  train.py:96-99     main.py:67-70
  ResBlock                    train.py:112-122   main.py:83-90
  deep tower / final linear   train.py:161-170   main.py:120-127
  BCE-with-logits mean loss   train.py:206,224
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm1d default, train.py:106
BN_MOMENTUM = 0.1    # nn.BatchNorm1d default


# --------------------------------------------------------------------------- shapes
def cat_widths(cat_dims: Dict[str, int]) -> List[int]:
    """Embedding width of each categorical table: int(sqrt(n)) + 1  (train.py:139)."""
    return [int(np.sqrt(n)) + 1 for n in cat_dims.values()]


def input_dim(emb_dim: int, cat_dims: Dict[str, int], n_num: int) -> int:
    """D = 2E + sum(c_i) + n_num  (train.py:140-141)."""
    return 2 * emb_dim + sum(cat_widths(cat_dims)) + n_num


def make_state(n_users: int, n_items: int, cat_dims: Dict[str, int], n_num: int, params: dict,
               seed: int = 0, emb_scale: float = 1.0, randomize_bn: bool = False,
               dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """A state_dict with the reference's key names / shapes.

    Initial distributions follow torch defaults (N(0,1) embeddings, U(+-1/sqrt(fan_in))
    linears, BN affine 1/0) but NOT torch's RNG stream -- golden tests load real
    reference state_dicts instead.  ``randomize_bn`` gives the "trained-like" variant
    of SURVEY.md section 8d.
    """
    g = torch.Generator().manual_seed(seed)
    E, H = params["emb_dim"], params["hidden_dim"]
    L, R = params["n_cross_layers"], params.get("n_res_blocks", 2)
    D = input_dim(E, cat_dims, n_num)

    def uni(shape, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["user_embedding.weight"] = torch.randn(n_users, E, generator=g, dtype=torch.float64) * emb_scale
    sd["item_embedding.weight"] = torch.randn(n_items, E, generator=g, dtype=torch.float64) * emb_scale
    for i, (n, w) in enumerate(zip(cat_dims.values(), cat_widths(cat_dims))):
        sd[f"cat_embeddings.{i}.weight"] = torch.randn(n, w, generator=g, dtype=torch.float64) * emb_scale
    sd["initial_deep_layer.weight"] = uni((H, D), D)
    sd["initial_deep_layer.bias"] = uni((H,), D)
    for r in range(R):
        for j in (1, 2):
            sd[f"res_blocks.{r}.layer{j}.weight"] = uni((H, H), H)
            sd[f"res_blocks.{r}.layer{j}.bias"] = uni((H,), H)
        for j in (1, 2):
            if randomize_bn:
                sd[f"res_blocks.{r}.bn{j}.weight"] = 0.5 + torch.rand(H, generator=g, dtype=torch.float64)
                sd[f"res_blocks.{r}.bn{j}.bias"] = torch.randn(H, generator=g, dtype=torch.float64) * 0.2
                sd[f"res_blocks.{r}.bn{j}.running_mean"] = torch.randn(H, generator=g, dtype=torch.float64) * 0.3
                sd[f"res_blocks.{r}.bn{j}.running_var"] = 0.5 + torch.rand(H, generator=g, dtype=torch.float64)
            else:
                sd[f"res_blocks.{r}.bn{j}.weight"] = torch.ones(H, dtype=torch.float64)
                sd[f"res_blocks.{r}.bn{j}.bias"] = torch.zeros(H, dtype=torch.float64)
                sd[f"res_blocks.{r}.bn{j}.running_mean"] = torch.zeros(H, dtype=torch.float64)
                sd[f"res_blocks.{r}.bn{j}.running_var"] = torch.ones(H, dtype=torch.float64)
            sd[f"res_blocks.{r}.bn{j}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    # key order inside a CrossLayer is (b, w.weight): b is registered first? No --
    # nn.Module lists parameters in assignment order: w (a submodule) comes after the
    # directly-registered Parameter b in state_dict(); order is irrelevant to loading.
    for l in range(L):
        sd[f"cross_network.{l}.b"] = torch.zeros(D, dtype=torch.float64)
        sd[f"cross_network.{l}.w.weight"] = uni((1, D), D)
    sd["final_linear.weight"] = uni((1, H + D), H + D)
    sd["final_linear.bias"] = uni((1,), H + D)
    out = OrderedDict()
    for k, v in sd.items():
        out[k] = v if v.dtype == torch.long else v.to(dtype)
    return out


def model_shape(state: Dict[str, torch.Tensor]) -> dict:
    """Recover (E, H, L, R, D, n_cat) from a state_dict's keys and shapes."""
    n_cat = sum(1 for k in state if k.startswith("cat_embeddings."))
    R = len({k.split(".")[1] for k in state if k.startswith("res_blocks.")})
    L = len({k.split(".")[1] for k in state if k.startswith("cross_network.")})
    H, D = state["initial_deep_layer.weight"].shape
    return dict(E=state["user_embedding.weight"].shape[1], H=H, D=D, L=L, R=R, n_cat=n_cat)


# --------------------------------------------------------------------------- torch functional
def gather_concat(state, user_ids, item_ids, cat_features, num_features):
    """x0 = [U[u] | I[i] | C0[c0] | C1[c1] ... | num]   (train.py:156-159)."""
    n_cat = model_shape(state)["n_cat"]
    parts = [F.embedding(user_ids, state["user_embedding.weight"]),
             F.embedding(item_ids, state["item_embedding.weight"])]
    for i in range(n_cat):
        parts.append(F.embedding(cat_features[:, i], state[f"cat_embeddings.{i}.weight"]))
    parts.append(num_features.to(parts[0].dtype))
    return torch.cat(parts, dim=1)


def cross_layer(x, w, b):
    """y = x + x * (x . w) + b -- what train.py:96-99 computes (rank-1, uses the layer's own input)."""
    s = x @ w.reshape(-1, 1)                 # [B,1]
    return x + x * s + b


def batchnorm_train(z, gamma, beta, eps=BN_EPS):
    """Batch-statistic BatchNorm1d: biased variance for normalisation (train.py:115,119)."""
    mean = z.mean(dim=0)
    var = z.var(dim=0, unbiased=False)
    return (z - mean) * torch.rsqrt(var + eps) * gamma + beta, mean, var


def forward(state, user_ids, item_ids, cat_features, num_features, *, training: bool,
            drop_masks: Optional[Sequence[torch.Tensor]] = None, dropout_p: float = 0.0,
            update_running: bool = False, return_parts: bool = False):
    """Functional DCN_RecSys.forward (train.py:155-170).

    ``drop_masks[r]`` is a {0,1} keep-mask [B,H] for ResBlock r's dropout
    (train.py:117); the kept activations are scaled by 1/(1-p).  With ``None`` dropout is
    the identity (p=0 / eval), which is what every parity test uses.
    Returns logits of shape [B] (0-d when B == 1, matching ``.squeeze()`` at train.py:170).
    """
    shp = model_shape(state)
    x0 = gather_concat(state, user_ids, item_ids, cat_features, num_features)
    B = x0.shape[0]
    if training and B == 1:
        raise ValueError("Expected more than 1 value per channel when training")  # torch BN behaviour
    h = F.linear(x0, state["initial_deep_layer.weight"], state["initial_deep_layer.bias"])
    for r in range(shp["R"]):
        p = f"res_blocks.{r}."
        identity = h
        outs = []
        t = h
        for j in (1, 2):
            z = F.linear(t, state[p + f"layer{j}.weight"], state[p + f"layer{j}.bias"])
            if training:
                y, mean, var = batchnorm_train(z, state[p + f"bn{j}.weight"], state[p + f"bn{j}.bias"])
                if update_running:
                    with torch.no_grad():
                        n = z.shape[0]
                        state[p + f"bn{j}.running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                        state[p + f"bn{j}.running_var"].mul_(1 - BN_MOMENTUM).add_(
                            BN_MOMENTUM * var * (n / (n - 1)))
                        state[p + f"bn{j}.num_batches_tracked"].add_(1)
            else:
                rm, rv = state[p + f"bn{j}.running_mean"], state[p + f"bn{j}.running_var"]
                y = (z - rm) * torch.rsqrt(rv + BN_EPS) * state[p + f"bn{j}.weight"] + state[p + f"bn{j}.bias"]
            if j == 1:
                t = torch.relu(y)
                if training and drop_masks is not None:
                    t = t * drop_masks[r].to(t.dtype) * (1.0 / (1.0 - dropout_p))
            else:
                t = torch.relu(y + identity)
            outs.append(t)
        h = t
    c = x0
    for l in range(shp["L"]):
        c = cross_layer(c, state[f"cross_network.{l}.w.weight"], state[f"cross_network.{l}.b"])
    wf = state["final_linear.weight"]
    # deep part FIRST in the concat (train.py:169)
    logits = (h @ wf[0, :shp["H"]] + c @ wf[0, shp["H"]:] + state["final_linear.bias"][0])
    logits = logits.squeeze()
    if return_parts:
        return logits, dict(x0=x0, deep=h, cross=c)
    return logits


def bce_with_logits_mean(logits, y):
    """nn.BCEWithLogitsLoss() default reduction='mean'  (train.py:206,224)."""
    return (F.softplus(logits) - y * logits).mean()


def forward_backward(state, user_ids, item_ids, cat_features, num_features, *,
                     grad_logits: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                     drop_masks=None, dropout_p: float = 0.0, dtype=None):
    """Train-mode forward + autograd backward.  Returns (logits, grads, loss_or_None).

    Either an explicit upstream ``grad_logits`` (kink-masked parity tests) or
    ``labels`` (then the loss is BCE-with-logits mean, train.py:224-225).
    Embedding-table gradients are DENSE, like nn.Embedding(sparse=False).
    """
    leaves = OrderedDict()
    for k, v in state.items():
        if v.dtype.is_floating_point and "running_" not in k:
            t = v.detach().clone()
            if dtype is not None:
                t = t.to(dtype)
            leaves[k] = t.requires_grad_(True)
        else:
            leaves[k] = v.detach().clone() if dtype is None or not v.dtype.is_floating_point else v.detach().to(dtype)
    if dtype is not None:
        num_features = num_features.to(dtype)
    logits = forward(leaves, user_ids, item_ids, cat_features, num_features, training=True,
                     drop_masks=drop_masks, dropout_p=dropout_p)
    loss = None
    if grad_logits is not None:
        logits.backward(gradient=grad_logits.to(logits.dtype).reshape(logits.shape))
    else:
        loss = bce_with_logits_mean(logits, labels.to(logits.dtype))
        loss.backward()
    grads = OrderedDict((k, v.grad) for k, v in leaves.items() if v.requires_grad)
    return logits.detach(), grads, (None if loss is None else loss.detach())


# --------------------------------------------------------------------------- numpy closed form
def _np(state, k):
    return state[k].detach().double().numpy()


def np_cross_fwd(x, w, b):
    """One CrossLayer: s = x.w ; y = x*(1+s) + b   (SURVEY section 4 KAT)."""
    s = x @ w
    return x * (1.0 + s)[:, None] + b, s


def np_cross_bwd(x, w, g):
    """gx = g(1+s) + w (g.x) ; gw = sum_rows (g.x) x ; gb = sum_rows g."""
    s = x @ w
    gx_dot = np.einsum("bd,bd->b", g, x)
    gx = g * (1.0 + s)[:, None] + gx_dot[:, None] * w[None, :]
    gw = (gx_dot[:, None] * x).sum(0)
    gb = g.sum(0)
    return gx, gw, gb


def np_forward_backward(state, user_ids, item_ids, cat_features, num_features, grad_logits,
                        drop_masks=None, dropout_p: float = 0.0):
    """float64 closed-form train-mode forward and hand-derived backward.

    This is the algorithm of the CUDA path, step for step:
      fwd : gather+concat -> [GEMM+bias -> batch stats -> normalise/ReLU(/mask/residual)]* ->
            cross layers in registers -> dual dot product for the logit
      bwd : dlogit (x) w_f -> per ResBlock (reverse): ReLU mask from the saved output,
            BN backward reduce (sum dy, sum dy*xhat) -> BN backward apply -> wgrad / dgrad
            -> cross backward with forward recomputation from x0 -> sorted-segment
            embedding scatter.
    Returns (logits, grads) with the state_dict's key names.
    """
    shp = model_shape(state)
    H, D, L, R, n_cat = shp["H"], shp["D"], shp["L"], shp["R"], shp["n_cat"]
    u = user_ids.numpy(); it = item_ids.numpy(); cf = cat_features.numpy()
    U = _np(state, "user_embedding.weight"); I = _np(state, "item_embedding.weight")
    C = [_np(state, f"cat_embeddings.{i}.weight") for i in range(n_cat)]
    x0 = np.concatenate([U[u], I[it]] + [C[i][cf[:, i]] for i in range(n_cat)] +
                        [num_features.double().numpy()], axis=1)
    B = x0.shape[0]
    scale = 1.0 / (1.0 - dropout_p)
    W0 = _np(state, "initial_deep_layer.weight"); b0 = _np(state, "initial_deep_layer.bias")
    h = x0 @ W0.T + b0
    saved = []
    for r in range(R):
        p = f"res_blocks.{r}."
        W1 = _np(state, p + "layer1.weight"); W2 = _np(state, p + "layer2.weight")
        z1 = h @ W1.T + _np(state, p + "layer1.bias")
        m1 = z1.mean(0); v1 = z1.var(0); rs1 = 1.0 / np.sqrt(v1 + BN_EPS); xh1 = (z1 - m1) * rs1
        a1 = np.maximum(xh1 * _np(state, p + "bn1.weight") + _np(state, p + "bn1.bias"), 0.0)
        d1 = a1 * (drop_masks[r].double().numpy() * scale) if drop_masks is not None else a1
        z2 = d1 @ W2.T + _np(state, p + "layer2.bias")
        m2 = z2.mean(0); v2 = z2.var(0); rs2 = 1.0 / np.sqrt(v2 + BN_EPS); xh2 = (z2 - m2) * rs2
        out = np.maximum(xh2 * _np(state, p + "bn2.weight") + _np(state, p + "bn2.bias") + h, 0.0)
        saved.append(dict(h_in=h, xh1=xh1, rs1=rs1, d1=d1, xh2=xh2, rs2=rs2, out=out, W1=W1, W2=W2))
        h = out
    cs = [x0]
    for l in range(L):
        y, _ = np_cross_fwd(cs[-1], _np(state, f"cross_network.{l}.w.weight")[0], _np(state, f"cross_network.{l}.b"))
        cs.append(y)
    wf = _np(state, "final_linear.weight")[0]
    logits = h @ wf[:H] + cs[-1] @ wf[H:] + _np(state, "final_linear.bias")[0]

    g = grad_logits.double().numpy().reshape(B)
    grads: Dict[str, np.ndarray] = {}
    grads["final_linear.bias"] = np.array([g.sum()])
    grads["final_linear.weight"] = np.concatenate([g @ h, g @ cs[-1]])[None, :]
    dh = g[:, None] * wf[None, :H]
    for r in reversed(range(R)):
        p = f"res_blocks.{r}."; s = saved[r]
        dy2 = dh * (s["out"] > 0)                      # ReLU after the residual add (train.py:120-121)
        gam2 = _np(state, p + "bn2.weight")
        grads[p + "bn2.bias"] = dy2.sum(0); grads[p + "bn2.weight"] = (dy2 * s["xh2"]).sum(0)
        dz2 = gam2 * s["rs2"] * (dy2 - grads[p + "bn2.bias"] / B - s["xh2"] * grads[p + "bn2.weight"] / B)
        grads[p + "layer2.bias"] = dz2.sum(0); grads[p + "layer2.weight"] = dz2.T @ s["d1"]
        dd1 = dz2 @ s["W2"]
        dy1 = dd1 * scale * (s["d1"] > 0)               # ReLU and dropout masks from the saved output
        gam1 = _np(state, p + "bn1.weight")
        grads[p + "bn1.bias"] = dy1.sum(0); grads[p + "bn1.weight"] = (dy1 * s["xh1"]).sum(0)
        dz1 = gam1 * s["rs1"] * (dy1 - grads[p + "bn1.bias"] / B - s["xh1"] * grads[p + "bn1.weight"] / B)
        grads[p + "layer1.bias"] = dz1.sum(0); grads[p + "layer1.weight"] = dz1.T @ s["h_in"]
        dh = dz1 @ s["W1"] + dy2                        # identity path
    grads["initial_deep_layer.bias"] = dh.sum(0)
    grads["initial_deep_layer.weight"] = dh.T @ x0
    dx0 = dh @ W0
    gc = g[:, None] * wf[None, H:]
    for l in reversed(range(L)):
        gc, gw, gb = np_cross_bwd(cs[l], _np(state, f"cross_network.{l}.w.weight")[0], gc)
        grads[f"cross_network.{l}.w.weight"] = gw[None, :]; grads[f"cross_network.{l}.b"] = gb
    dx0 = dx0 + gc
    E = U.shape[1]
    grads["user_embedding.weight"] = segment_scatter(u, dx0[:, :E], U.shape[0])
    grads["item_embedding.weight"] = segment_scatter(it, dx0[:, E:2 * E], I.shape[0])
    off = 2 * E
    for i in range(n_cat):
        w = C[i].shape[1]
        grads[f"cat_embeddings.{i}.weight"] = segment_scatter(cf[:, i], dx0[:, off:off + w], C[i].shape[0])
        off += w
    return logits, grads


def segment_scatter(ids: np.ndarray, g: np.ndarray, n_rows: int) -> np.ndarray:
    """Dense embedding gradient by stable sort + per-segment sum in batch order
    (the deterministic order of the CUDA scatter; same result as embedding_dense_backward)."""
    order = np.argsort(ids, kind="stable")
    out = np.zeros((n_rows, g.shape[1]), dtype=g.dtype)
    sid = ids[order]
    starts = np.flatnonzero(np.r_[True, sid[1:] != sid[:-1]])
    ends = np.r_[starts[1:], len(sid)]
    for s, e in zip(starts, ends):
        acc = np.zeros(g.shape[1], dtype=g.dtype)
        for p in range(s, e):
            acc = acc + g[order[p]]
        out[sid[s]] = acc
    return out


# --------------------------------------------------------------------------- parity metrics
def max_abs_normalised(a, b) -> float:
    """max|a-b| / max|b|  -- the logit / gradient parity metric of SURVEY.md section 8d."""
    a = torch.as_tensor(a).double().reshape(-1); b = torch.as_tensor(b).double().reshape(-1)
    denom = float(b.abs().max()) if b.numel() else 0.0
    if denom == 0.0:
        return float((a - b).abs().max()) if a.numel() else 0.0
    return float((a - b).abs().max()) / denom


def relu_preactivations(state, user_ids, item_ids, cat_features, num_features, dtype=torch.float64):
    """Train-mode inputs of every ReLU, in order [blk0.relu1, blk0.relu2, blk1.relu1, ...], each [B,H]."""
    shp = model_shape(state)
    st = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in state.items()}
    x0 = gather_concat(st, user_ids, item_ids, cat_features, num_features.to(dtype))
    h = F.linear(x0, st["initial_deep_layer.weight"], st["initial_deep_layer.bias"])
    pre = []
    for r in range(shp["R"]):
        p = f"res_blocks.{r}."
        z1 = F.linear(h, st[p + "layer1.weight"], st[p + "layer1.bias"])
        y1, _, _ = batchnorm_train(z1, st[p + "bn1.weight"], st[p + "bn1.bias"])
        z2 = F.linear(torch.relu(y1), st[p + "layer2.weight"], st[p + "layer2.bias"])
        y2, _, _ = batchnorm_train(z2, st[p + "bn2.weight"], st[p + "bn2.bias"])
        pre += [y1, y2 + h]
        h = torch.relu(y2 + h)
    return pre


def desensitize_relus(state, user_ids, item_ids, cat_features, num_features, search: float = 0.05):
    """Returns (new_state, margin): a copy of ``state`` whose BatchNorm biases are nudged (by at most
    ``search`` per channel) so that, for THIS batch, no train-mode ReLU input lies within ``margin``
    of zero.  Gradient-parity tests use it to compare arithmetic instead of ReLU discontinuities:
    zeroing a near-kink row's upstream gradient is not enough in train mode because BatchNorm
    couples the rows (a flipped ReLU in a masked row still leaks through the batch statistics), and
    dropping rows only moves the problem (the statistics change and new near-kinks appear).
    For each ReLU layer in forward order and each channel, the bias shift puts zero in the middle of
    the widest gap between consecutive sorted pre-activations near zero."""
    st = OrderedDict((k, v.clone()) for k, v in state.items())
    shp = model_shape(st)
    names = []
    for r in range(shp["R"]):
        names += [f"res_blocks.{r}.bn1.bias", f"res_blocks.{r}.bn2.bias"]
    margin = float("inf")
    for j, name in enumerate(names):
        y = relu_preactivations(st, user_ids, item_ids, cat_features, num_features)[j]      # [B,H] float64
        v, _ = torch.sort(y, dim=0)
        mid = 0.5 * (v[1:] + v[:-1])
        gap = v[1:] - v[:-1]
        gap = torch.where(mid.abs() <= search, gap, torch.zeros_like(gap))
        best = gap.argmax(dim=0)                                                            # per channel
        cols = torch.arange(y.shape[1])
        shift = -mid[best, cols]
        shift = torch.where(gap[best, cols] > 0, shift, torch.zeros_like(shift))
        st[name] = (st[name].double() + shift).to(st[name].dtype)
    for y in relu_preactivations(st, user_ids, item_ids, cat_features, num_features):
        margin = min(margin, float(y.abs().min()))
    return st, margin


def kink_mask(state, user_ids, item_ids, cat_features, num_features, thresh: float = 1e-5) -> torch.Tensor:
    """Rows whose float64 forward has a ReLU pre-activation within ``thresh`` of zero
    (their upstream gradient is zeroed in gradient-parity tests)."""
    bad = torch.zeros(user_ids.shape[0], dtype=torch.bool)
    for y in relu_preactivations(state, user_ids, item_ids, cat_features, num_features):
        bad |= (y.abs() < thresh).any(dim=1)
    return bad
