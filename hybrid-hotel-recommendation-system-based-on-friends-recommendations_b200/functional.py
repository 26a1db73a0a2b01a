"""Operator-level autograd bindings over the C ABI (one stage of the DCN-R path each).

``DCN_RecSys`` itself dispatches to the whole-model entry points (model.py); these functions back
the stand-alone ``CrossLayer`` / ``ResBlock`` modules and the operator parity tests.  Everything
here requires CUDA tensors -- the library has no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
from torch.autograd import Function

from . import _cabi as C


def _scratch(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = C.ptr(t)
    return arr


def _prec(precision) -> int:
    return C.PRECISIONS[precision] if isinstance(precision, str) else int(precision)


# ------------------------------------------------------------------------------------------ cross network
class _CrossFn(Function):
    @staticmethod
    def forward(ctx, x, n_layers, *wb):
        C.require_cuda(x)
        x = _f32c(x)
        ws = [_f32c(w).reshape(-1) for w in wb[:n_layers]]
        bs = [_f32c(b) for b in wb[n_layers:]]
        B, D = x.shape
        y = torch.empty_like(x)
        C.check(C.lib().dcnr_cross_fwd(C.ptr(x), D, B, D, n_layers, _ptr_array(ws), _ptr_array(bs), C.ptr(y), D,
                                       C.stream()))
        ctx.save_for_backward(x, *ws, *bs)
        ctx.n_layers = n_layers
        return y

    @staticmethod
    def backward(ctx, gy):
        L = ctx.n_layers
        x, ws, bs = ctx.saved_tensors[0], ctx.saved_tensors[1:1 + L], ctx.saved_tensors[1 + L:]
        gy = _f32c(gy)
        B, D = x.shape
        gx = torch.empty_like(x)
        gws = [torch.empty(D, device=x.device, dtype=torch.float32) for _ in range(L)]
        gbs = [torch.empty(D, device=x.device, dtype=torch.float32) for _ in range(L)]
        nbytes = C.lib().dcnr_cross_bwd_scratch_bytes(B, D, L)
        ws_buf = _scratch(nbytes, x.device)
        C.check(C.lib().dcnr_cross_bwd(C.ptr(x), D, B, D, L, _ptr_array(ws), _ptr_array(bs), C.ptr(gy), D, C.ptr(gx), D,
                                       _ptr_array(gws), _ptr_array(gbs), C.ptr(ws_buf), ws_buf.numel(), C.stream()))
        return (gx, None) + tuple(g.reshape(1, D) for g in gws) + tuple(gbs)


def cross_network(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> torch.Tensor:
    """All CrossLayers in one kernel: y <- y * (1 + y . w_l) + b_l  (train.py:96-99, :167-168).

    ``weights[l]`` is ``cross_network.l.w.weight`` ([1, D]); ``biases[l]`` is ``cross_network.l.b`` ([D])."""
    return _CrossFn.apply(x, len(weights), *weights, *biases)


# ------------------------------------------------------------------------------------------ DCN-v2 cross (opt-in)
def _pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    t = _f32c(t)
    return t if t.shape[-1] == width else torch.nn.functional.pad(t, (0, width - t.shape[-1]))


class _CrossV2Fn(Function):
    """All layers of the full-matrix cross network on rows padded to a multiple of 32 floats:
    x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l.  Saves the layer inputs; recomputes u_l = x_l W_l^T + b_l in backward."""

    @staticmethod
    def forward(ctx, x0, n_layers, precision, *wb):
        C.require_cuda(x0)
        B, D = x0.shape
        Dp = (D + 31) // 32 * 32
        p = _prec(precision)
        x0p = _pad_cols(x0, Dp)
        ws = [torch.nn.functional.pad(_f32c(w), (0, Dp - D, 0, Dp - D)) for w in wb[:n_layers]]
        bs = [_pad_cols(b, Dp) for b in wb[n_layers:]]
        xs = [x0p]
        lws = _scratch(C.lib().dcnr_linear_workspace_bytes(Dp, Dp, p), x0p.device)
        for l in range(n_layers):
            y = torch.empty_like(x0p)
            C.check(C.lib().dcnr_cross_v2_fwd(C.ptr(x0p), Dp, C.ptr(xs[-1]), Dp, C.ptr(ws[l]), Dp, C.ptr(bs[l]), C.ptr(y), Dp,
                                              B, Dp, p, C.ptr(lws), lws.numel(), C.stream()))
            xs.append(y)
        ctx.save_for_backward(*xs[:-1], *ws, *bs)
        ctx.meta = (n_layers, D, Dp, p)
        return xs[-1][:, :D].contiguous() if Dp != D else xs[-1]

    @staticmethod
    def backward(ctx, gy):
        L, D, Dp, p = ctx.meta
        xs, ws, bs = ctx.saved_tensors[:L], ctx.saved_tensors[L:2 * L], ctx.saved_tensors[2 * L:]
        x0p = xs[0]
        B = x0p.shape[0]
        lib, st = C.lib(), C.stream()
        g = _pad_cols(gy, Dp)
        dx0 = torch.zeros_like(x0p)
        gws, gbs = [None] * L, [None] * L
        scratch = _scratch(lib.dcnr_linear_wgrad_scratch_bytes(B, Dp, Dp), x0p.device)
        lws = _scratch(lib.dcnr_linear_workspace_bytes(Dp, Dp, p), x0p.device)
        u, gm = torch.empty_like(x0p), torch.empty_like(x0p)
        for l in reversed(range(L)):
            # u = x_l W^T + b ; gm = g * x0 ; dx0 += g * u
            C.check(lib.dcnr_linear_fwd(C.ptr(xs[l]), Dp, C.ptr(ws[l]), Dp, C.ptr(bs[l]), None, None, 0, 0, C.ptr(u), Dp, B, Dp,
                                        Dp, p, C.ptr(lws), lws.numel(), st))
            C.check(lib.dcnr_cross_v2_bwd_prep(C.ptr(g), Dp, C.ptr(x0p), Dp, C.ptr(u), Dp, C.ptr(gm), Dp, C.ptr(dx0), Dp, 1, B,
                                               Dp, st))
            gw = torch.empty((Dp, Dp), device=x0p.device, dtype=torch.float32)
            gb = torch.empty(Dp, device=x0p.device, dtype=torch.float32)
            C.check(lib.dcnr_linear_wgrad(C.ptr(gm), Dp, C.ptr(xs[l]), Dp, C.ptr(gw), Dp, C.ptr(gb), B, Dp, Dp, p,
                                          C.ptr(scratch), scratch.numel(), st))
            gws[l], gbs[l] = gw[:D, :D].contiguous(), gb[:D].contiguous()
            g_next = torch.empty_like(x0p)       # dx_l = gm W + g
            C.check(lib.dcnr_linear_dgrad(C.ptr(gm), Dp, C.ptr(ws[l]), Dp, C.ptr(g), Dp, C.ptr(g_next), Dp, B, Dp, Dp, p,
                                          C.ptr(lws), lws.numel(), st))
            g = g_next
        dx0 += g                                  # x_0 is also the first layer's input
        gx0 = dx0[:, :D].contiguous() if Dp != D else dx0
        return (gx0, None, None) + tuple(gws) + tuple(gbs)


def cross_network_v2(x0: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                     precision="tf32x3") -> torch.Tensor:
    """DCN-v2 cross network (opt-in): x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l with W_l [D, D].
    Each layer is one tcgen05 GEMM with bias / Hadamard / residual in the epilogue (dcnr_cross_v2_fwd)."""
    return _CrossV2Fn.apply(x0, len(weights), precision, *weights, *biases)


# ------------------------------------------------------------------------------------------ linear
def linear_forward_raw(x, w, bias=None, col_scale=None, residual=None, relu=False, precision="fp32"):
    C.require_cuda(x, w)
    x, w = _f32c(x), _f32c(w)
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty((m, n), device=x.device, dtype=torch.float32)
    residual = None if residual is None else _f32c(residual)
    lws = _scratch(C.lib().dcnr_linear_workspace_bytes(n, k, _prec(precision)), x.device)
    C.check(C.lib().dcnr_linear_fwd(C.ptr(x), k, C.ptr(w), k, C.ptr(bias), C.ptr(col_scale), C.ptr(residual), n,
                                    1 if relu else 0, C.ptr(y), n, m, n, k, _prec(precision), C.ptr(lws), lws.numel(), C.stream()))
    return y


class _LinearFn(Function):
    @staticmethod
    def forward(ctx, x, w, bias, precision):
        y = linear_forward_raw(x, w, None if bias is None else _f32c(bias), precision=precision)
        ctx.save_for_backward(_f32c(x), _f32c(w))
        ctx.has_bias = bias is not None
        ctx.precision = precision
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _f32c(gy)
        m, k = x.shape
        n = w.shape[0]
        p = _prec(ctx.precision)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            lws = _scratch(C.lib().dcnr_linear_workspace_bytes(n, k, p), x.device)
            C.check(C.lib().dcnr_linear_dgrad(C.ptr(gy), n, C.ptr(w), k, None, 0, C.ptr(gx), k, m, n, k, p, C.ptr(lws), lws.numel(),
                                              C.stream()))
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(w)
            gb = torch.empty(n, device=x.device, dtype=torch.float32) if ctx.has_bias else None
            ws = _scratch(C.lib().dcnr_linear_wgrad_scratch_bytes(m, n, k), x.device)
            C.check(C.lib().dcnr_linear_wgrad(C.ptr(gy), n, C.ptr(x), k, C.ptr(gw), k, C.ptr(gb), m, n, k, p,
                                              C.ptr(ws), ws.numel(), C.stream()))
        return gx, gw, gb, None


def linear(x, weight, bias=None, precision="fp32"):
    """y = x W^T + b  (nn.Linear; train.py:143, :105, :109)."""
    return _LinearFn.apply(x, weight, bias, precision)


# ------------------------------------------------------------------------------------------ BN + ReLU (+dropout, +residual)
class _BNActFn(Function):
    @staticmethod
    def forward(ctx, z, gamma, beta, residual, running_mean, running_var, num_batches_tracked, eps, momentum,
                drop_p, keep_mask, seed, layer_tag):
        C.require_cuda(z)
        z = _f32c(z)
        m, n = z.shape
        if m < 2:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
        mean = torch.empty(n, device=z.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        ws = _scratch(C.lib().dcnr_bn_scratch_bytes(m, n), z.device)
        C.check(C.lib().dcnr_bn_stats(C.ptr(z), n, m, n, eps, momentum, C.ptr(mean), C.ptr(rstd), C.ptr(running_mean),
                                      C.ptr(running_var), C.ptr(num_batches_tracked), C.ptr(ws), ws.numel(), C.stream()))
        C.mark_mutated([t for t in (running_mean, running_var, num_batches_tracked) if t is not None])
        residual = None if residual is None else _f32c(residual)
        out = torch.empty_like(z)
        gamma, beta = _f32c(gamma), _f32c(beta)
        C.check(C.lib().dcnr_bn_act_fwd(C.ptr(z), n, C.ptr(mean), C.ptr(rstd), C.ptr(gamma), C.ptr(beta), C.ptr(residual),
                                        n, C.ptr(keep_mask), drop_p, seed, layer_tag, C.ptr(out), n, m, n, C.stream()))
        ctx.save_for_backward(z, out, mean, rstd, gamma)
        ctx.post_scale = 1.0 / (1.0 - drop_p) if (drop_p > 0 or keep_mask is not None) else 1.0
        ctx.has_residual = residual is not None
        return out

    @staticmethod
    def backward(ctx, g):
        z, out, mean, rstd, gamma = ctx.saved_tensors
        g = _f32c(g)
        m, n = z.shape
        dz = torch.empty_like(z)
        dy = torch.empty_like(z) if ctx.has_residual else None
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        ws = _scratch(C.lib().dcnr_bn_scratch_bytes(m, n), z.device)
        C.check(C.lib().dcnr_bn_act_bwd(C.ptr(g), n, C.ptr(out), n, C.ptr(z), n, C.ptr(mean), C.ptr(rstd), C.ptr(gamma),
                                        ctx.post_scale, C.ptr(dz), n, C.ptr(dy), n, C.ptr(dgamma), C.ptr(dbeta), None,
                                        m, n, C.ptr(ws), ws.numel(), C.stream()))
        return (dz, dgamma, dbeta, dy) + (None,) * 9


def batchnorm_relu_train(z, gamma, beta, residual=None, running_mean=None, running_var=None,
                         num_batches_tracked=None, eps=1e-5, momentum=0.1, drop_p=0.0, keep_mask=None, seed=0,
                         layer_tag=0):
    """relu(BatchNorm1d_train(z) + residual), then dropout  (train.py:115-117 and :119-121)."""
    return _BNActFn.apply(z, gamma, beta, residual, running_mean, running_var, num_batches_tracked, float(eps),
                          float(momentum), float(drop_p), keep_mask, int(seed), int(layer_tag))


def fold_batchnorm(bn_weight, bn_bias, running_mean, running_var, lin_bias, eps=1e-5):
    """Eval-mode BatchNorm folded onto a preceding linear: (scale, shift) with
    BN(acc + b) = acc*scale + shift.  Folded in float64 (6.4e-7 vs reference)."""
    s = bn_weight.double() / torch.sqrt(running_var.double() + eps)
    shift = bn_bias.double() + (lin_bias.double() - running_mean.double()) * s
    return s.float().contiguous(), shift.float().contiguous()


# ------------------------------------------------------------------------------------------ misc raw ops
def rowdot(a, w, extra=None, bias=None):
    C.require_cuda(a)
    a = _f32c(a)
    m, n = a.shape
    out = torch.empty(m, device=a.device, dtype=torch.float32)
    C.check(C.lib().dcnr_rowdot_fwd(C.ptr(a), n, C.ptr(_f32c(w)), C.ptr(extra), C.ptr(bias), C.ptr(out), m, n,
                                    C.stream()))
    return out


def bce_with_logits(logits, labels, want_grad=True):
    """nn.BCEWithLogitsLoss() mean loss and dL/dlogits in one pass (train.py:206,224)."""
    C.require_cuda(logits, labels)
    logits, labels = _f32c(logits).reshape(-1), _f32c(labels).reshape(-1)
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits) if want_grad else None
    scratch = torch.empty(4096, device=logits.device, dtype=torch.float32)
    C.check(C.lib().dcnr_bce_with_logits(C.ptr(logits), C.ptr(labels), logits.numel(), C.ptr(loss), C.ptr(grad),
                                         C.ptr(scratch), C.stream()))
    return loss, grad


def adam_step_(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
               decoupled=False):
    """In-place fused Adam/AdamW update of one dense tensor (torch.optim semantics; train.py:201-204,226)."""
    C.require_cuda(param, grad, exp_avg, exp_avg_sq)
    assert param.is_contiguous() and grad.is_contiguous() and exp_avg.is_contiguous() and exp_avg_sq.is_contiguous()
    C.check(C.lib().dcnr_adam_step(C.ptr(param), C.ptr(grad), C.ptr(exp_avg), C.ptr(exp_avg_sq), param.numel(), lr,
                                   betas[0], betas[1], eps, weight_decay, 1 if decoupled else 0, int(step), C.stream()))
    C.mark_mutated([param, exp_avg, exp_avg_sq])
