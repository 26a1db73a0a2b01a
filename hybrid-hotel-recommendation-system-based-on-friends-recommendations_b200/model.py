"""Drop-in ``CrossLayer`` / ``ResBlock`` / ``DCN_RecSys`` for the reference's model classes
(train.py:90-170, duplicated at main.py:61-127), computing on B200 through libdcnr_sm100a.so.

Same constructor signatures, same ``forward`` signature and output-shape rule, same parameter
names and shapes (a reference ``final_dcn_model.pth`` loads with ``load_state_dict`` unchanged,
stock ``torch.optim`` optimizers work on ``.parameters()``), same train()/eval() behaviour
(batch-statistic BatchNorm that updates running stats and ``num_batches_tracked``, dropout only in
train()).  The parameter *containers* are ordinary ``nn.Embedding`` / ``nn.Linear`` /
``nn.BatchNorm1d`` objects so that key names match; their own ``forward`` methods are never called.

There is no CPU path: inputs and parameters must be CUDA tensors.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _cabi as C
from . import functional as F_


class CrossLayer(nn.Module):
    """y = x + x * (x . w) + b   (train.py:90-99 / main.py:61-70: rank-1, uses the layer's own input)."""

    def __init__(self, input_dim: int):
        super().__init__()
        self.w = nn.Linear(input_dim, 1, bias=False)
        self.b = nn.Parameter(torch.zeros(input_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F_.cross_network(x, [self.w.weight], [self.b])


class CrossLayerV2(nn.Module):
    """OPT-IN DCN-v2 cross layer y = x0 * (W x + b) + x with a full [D, D] weight (BASELINE.json north_star item 2,
    SURVEY 8f-4).  Not in the reference (its CrossLayer above is rank-1) -- different parameters, own oracle
    (oracle/cross_v2_oracle.py), never the default."""

    def __init__(self, input_dim: int, precision: str = "tf32x3"):
        super().__init__()
        self.w = nn.Linear(input_dim, input_dim, bias=False)
        self.b = nn.Parameter(torch.zeros(input_dim))
        self.precision = precision

    def forward(self, x0: torch.Tensor, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x is not x0:
            raise ValueError("stand-alone CrossLayerV2 takes x == x0; stack layers with CrossNetworkV2")
        return F_.cross_network_v2(x0, [self.w.weight], [self.b], self.precision)


class CrossNetworkV2(nn.Module):
    """n_layers CrossLayerV2 applied from x0 (x_{l+1} = x0 * (W_l x_l + b_l) + x_l), one fused GEMM per layer."""

    def __init__(self, input_dim: int, n_layers: int, precision: str = "tf32x3"):
        super().__init__()
        self.layers = nn.ModuleList([CrossLayerV2(input_dim, precision) for _ in range(n_layers)])
        self.precision = precision

    def forward(self, x0: torch.Tensor) -> torch.Tensor:
        return F_.cross_network_v2(x0, [l.w.weight for l in self.layers], [l.b for l in self.layers], self.precision)


class ResBlock(nn.Module):
    """relu(BN2(L2(drop(relu(BN1(L1(x)))))) + x)   (train.py:102-122 / main.py:73-90)."""

    def __init__(self, hidden_dim: int, dropout: float):
        super().__init__()
        self.layer1 = nn.Linear(hidden_dim, hidden_dim)
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.layer2 = nn.Linear(hidden_dim, hidden_dim)
        self.bn2 = nn.BatchNorm1d(hidden_dim)
        self.precision = "fp32"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        C.require_cuda(x)
        p = self.precision
        if self.training:
            seed = int(torch.empty((), dtype=torch.int64).random_().item()) if self.dropout.p > 0 else 0
            z1 = F_.linear(x, self.layer1.weight, self.layer1.bias, p)
            d1 = F_.batchnorm_relu_train(z1, self.bn1.weight, self.bn1.bias, None, self.bn1.running_mean,
                                         self.bn1.running_var, self.bn1.num_batches_tracked, self.bn1.eps,
                                         self.bn1.momentum, self.dropout.p, None, seed, 0)
            z2 = F_.linear(d1, self.layer2.weight, self.layer2.bias, p)
            return F_.batchnorm_relu_train(z2, self.bn2.weight, self.bn2.bias, x, self.bn2.running_mean,
                                           self.bn2.running_var, self.bn2.num_batches_tracked, self.bn2.eps,
                                           self.bn2.momentum, 0.0, None, 0, 0)
        with torch.no_grad():
            s1, t1 = F_.fold_batchnorm(self.bn1.weight, self.bn1.bias, self.bn1.running_mean, self.bn1.running_var,
                                       self.layer1.bias, self.bn1.eps)
            s2, t2 = F_.fold_batchnorm(self.bn2.weight, self.bn2.bias, self.bn2.running_mean, self.bn2.running_var,
                                       self.layer2.bias, self.bn2.eps)
            t = F_.linear_forward_raw(x, self.layer1.weight, t1, s1, None, True, p)
            return F_.linear_forward_raw(t, self.layer2.weight, t2, s2, x, True, p)


class _DCNTrainFn(Function):
    """DCN_RecSys.forward in train() with autograd: dcnr_forward_train / dcnr_backward."""

    @staticmethod
    def forward(ctx, model, user_ids, item_ids, cat_features, num_features, *params):
        dims, pstruct = model._dims(), model._param_struct()
        B = user_ids.numel()
        batch = C.Batch(C.ptr(user_ids), C.ptr(item_ids), C.ptr(cat_features), C.ptr(num_features), B)
        dev = user_ids.device
        nbytes = C.lib().dcnr_workspace_bytes(dims, B, 1)
        saved = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if dims.dropout_p > 0 else 0
        mask = model._inject_drop_masks
        C.check(C.lib().dcnr_forward_train(dims, pstruct, batch, seed, C.ptr(mask), C.ptr(logits), C.ptr(saved),
                                           saved.numel(), C.stream()))
        C.mark_mutated(model.buffers())       # the kernels updated the BatchNorm running statistics through raw pointers
        ctx.model, ctx.saved_ws, ctx.dims, ctx.pstruct = model, saved, dims, pstruct
        ctx.inputs = (user_ids, item_ids, cat_features, num_features)
        ctx.params = params
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        model, dims = ctx.model, ctx.dims
        user_ids, item_ids, cat_features, num_features = ctx.inputs
        B = user_ids.numel()
        batch = C.Batch(C.ptr(user_ids), C.ptr(item_ids), C.ptr(cat_features), C.ptr(num_features), B)
        grad_logits = grad_logits.contiguous().float()
        grads = [torch.empty_like(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[5:])]
        gstruct = model._grad_struct(grads)
        dev = user_ids.device
        scratch = torch.empty(C.lib().dcnr_workspace_bytes(dims, B, 2), dtype=torch.uint8, device=dev)
        C.check(C.lib().dcnr_backward(dims, ctx.pstruct, batch, C.ptr(grad_logits), C.ptr(ctx.saved_ws),
                                      ctx.saved_ws.numel(), gstruct, C.ptr(scratch), scratch.numel(), C.stream()))
        ctx.saved_ws = None
        return (None, None, None, None, None) + tuple(grads)


class DCN_RecSys(nn.Module):
    """Drop-in for the reference's ``DCN_RecSys`` (train.py:125-170 / main.py:93-127).

    ``params`` needs ``emb_dim``, ``hidden_dim``, ``n_cross_layers``, ``dropout`` and optionally
    ``n_res_blocks`` (default 2, train.py:134); other keys (lr, batch_size, ...) are ignored like in
    the reference.  ``precision`` selects the dense-layer arithmetic:

    * ``"fp16x3"`` (default) -- parity-grade tensor-core mode.  eval(): the fused tower kernel (tcgen05 kind::f16 on a
      3-term error-compensated fp16 split, activations resident in tensor memory; logits within 1e-5 of the reference);
      train(): the tf32x3 kernels below.  Shapes the fused tower does not take (hidden_dim != 256, > 4 ResBlocks) run
      tf32x3 in eval() too.  An eval batch whose activations leave the fp16 range is re-run on tf32x3 automatically.
    * ``"tf32x3"`` -- tcgen05 kind::tf32 3-term split, one GEMM launch per layer; parity-grade: logits <= 1e-5, every
      gradient tensor within max(1e-5, 2 x the reference's own fp32 noise) of the float64 oracle.
    * ``"fp32"`` -- CUDA-core IEEE fp32 FMA GEMMs (strict, slow).
    * ``"tf32"`` / ``"bf16"`` -- single-pass tensor-core modes with STATED tolerances, not parity (bf16: fused eval tower with
      bf16 operands; training runs tf32).
    """

    def __init__(self, n_users: int, n_items: int, cat_dims: Dict[str, int], n_num_features: int,
                 params: Dict[str, Any], precision: Optional[str] = None):
        super().__init__()
        emb_dim = params['emb_dim']
        hidden_dim = params['hidden_dim']
        n_cross_layers = params['n_cross_layers']
        dropout = params['dropout']
        n_res_blocks = params.get('n_res_blocks', 2)

        self.user_embedding = nn.Embedding(n_users, emb_dim)
        self.item_embedding = nn.Embedding(n_items, emb_dim)
        widths = [int(np.sqrt(n_cat)) + 1 for n_cat in cat_dims.values()]
        self.cat_embeddings = nn.ModuleList([nn.Embedding(n, w) for n, w in zip(cat_dims.values(), widths)])
        input_dim = emb_dim * 2 + sum(widths) + n_num_features
        self.initial_deep_layer = nn.Linear(input_dim, hidden_dim)
        self.res_blocks = nn.ModuleList([ResBlock(hidden_dim, dropout) for _ in range(n_res_blocks)])
        self.cross_network = nn.ModuleList([CrossLayer(input_dim) for _ in range(n_cross_layers)])
        self.final_linear = nn.Linear(hidden_dim + input_dim, 1)

        if len(widths) > C.MAX_CAT or n_res_blocks > C.MAX_RES or n_cross_layers > C.MAX_CROSS:
            raise ValueError("model exceeds DCNR_MAX_CAT / DCNR_MAX_RES / DCNR_MAX_CROSS (8 each)")
        if hidden_dim % 4 != 0:
            raise ValueError("hidden_dim must be a multiple of 4 (the reference's search space uses multiples of 32)")
        self._shape = dict(emb_dim=emb_dim, n_num=n_num_features, hidden=hidden_dim, n_cross=n_cross_layers,
                           n_res=n_res_blocks, in_dim=input_dim, n_users=n_users, n_items=n_items,
                           cat_rows=list(cat_dims.values()), cat_width=widths, dropout=float(dropout))
        self.precision = precision or os.environ.get("DCNR_PRECISION", "fp16x3")
        self.check_ids = os.environ.get("DCNR_CHECK_IDS", "0") == "1"      # train(): bounds-check ids before every forward
        # eval(): the kernels report out-of-range ids and fp16-range overflow in a device flag word that is read back after
        # every forward (one 4-byte D2H).  A caller that batches many forwards (serving.RankingEngine) sets this to True and
        # calls check_eval_flags() once at the end.
        self.defer_eval_checks = False
        self._eval_flags = None
        self._tower_pack = None             # (cache key, device buffer) of the fused tower's prepared weights
        self._inject_drop_masks = None      # uint8 [n_res, B, H] keep-mask for parity tests

    # ---- C-ABI marshalling -----------------------------------------------------------------
    def _dims(self) -> C.Dims:
        s = self._shape
        d = C.Dims()
        d.emb_dim, d.n_cat, d.n_num, d.hidden = s["emb_dim"], len(s["cat_rows"]), s["n_num"], s["hidden"]
        d.n_cross, d.n_res, d.in_dim, d.in_dim_pad = s["n_cross"], s["n_res"], s["in_dim"], C.pad_dim(s["in_dim"])
        d.n_users, d.n_items = s["n_users"], s["n_items"]
        for i, (r, w) in enumerate(zip(s["cat_rows"], s["cat_width"])):
            d.cat_rows[i], d.cat_width[i] = r, w
        d.dropout_p = s["dropout"] if self.training else 0.0
        if self._eval_flags is not None:        # train-mode forwards record out-of-range ids here too; read by the next
            d.eval_flags = C.ptr(self._eval_flags)      # eval() forward or by check_eval_flags() (e.g. once per epoch)
        blk = self.res_blocks[0] if len(self.res_blocks) else None
        d.bn_eps = blk.bn1.eps if blk is not None else 1e-5
        d.bn_momentum = blk.bn1.momentum if blk is not None else 0.1
        d.precision = C.PRECISIONS[self.precision]
        step = getattr(self, "_dropout_step", None)        # device counter (training.GraphedTrainStep)
        d.dropout_step = C.ptr(step) if step is not None else None
        comm = getattr(self, "_comm", None)
        d.comm = comm.handle if (comm is not None and comm.world > 1 and self.training) else None
        d.dp_sparse_tables = 1 if (d.comm and getattr(self, "dp_sparse_embedding_grads", True) and
                                   getattr(self, "_row_override", None) is None) else 0
        d.dp_batch_cap = getattr(self, "_dp_batch_cap", 0) if d.dp_sparse_tables else 0
        rows = getattr(self, "_row_override", None)
        if rows is not None:                     # per-sample tables (row-sharded exchange): id = batch position
            d.n_users, d.n_items = rows[0].shape[0], rows[1].shape[0]
        return d

    def _ordered_params(self):
        """Parameters in the fixed order used by the autograd Function and dcnr_grads."""
        rows = getattr(self, "_row_override", None)
        tables = [self.user_embedding.weight, self.item_embedding.weight] if rows is None else list(rows)
        ps = tables + [e.weight for e in self.cat_embeddings]
        ps += [self.initial_deep_layer.weight, self.initial_deep_layer.bias]
        for blk in self.res_blocks:
            ps += [blk.layer1.weight, blk.layer1.bias, blk.bn1.weight, blk.bn1.bias,
                   blk.layer2.weight, blk.layer2.bias, blk.bn2.weight, blk.bn2.bias]
        for cl in self.cross_network:
            ps += [cl.w.weight, cl.b]
        ps += [self.final_linear.weight, self.final_linear.bias]
        return ps

    def _param_struct(self) -> C.Params:
        # the ctypes struct of 35 device pointers is rebuilt only when a storage moved (a single ranking request is ~100 us of
        # GPU time: marshalling must not cost as much)
        ps = self._ordered_params()
        bufs = []
        for blk in self.res_blocks:
            for bn in (blk.bn1, blk.bn2):
                bufs += [bn.running_mean, bn.running_var, bn.num_batches_tracked]
        key = tuple(t.data_ptr() for t in ps) + tuple(t.data_ptr() for t in bufs)
        cached = getattr(self, "_pstruct_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        p = self._build_param_struct(ps)
        self._pstruct_cache = (key, p)
        return p

    def _build_param_struct(self, ps) -> C.Params:
        p = C.Params()
        for t in ps:
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("DCN_RecSys parameters must be contiguous float32 CUDA tensors "
                                   "(call model.cuda(); there is no CPU path)")
        p.user_table, p.item_table = C.ptr(ps[0]), C.ptr(ps[1])
        for i, e in enumerate(self.cat_embeddings):
            p.cat_table[i] = C.ptr(e.weight)
        p.w0, p.b0 = C.ptr(self.initial_deep_layer.weight), C.ptr(self.initial_deep_layer.bias)
        for r, blk in enumerate(self.res_blocks):
            p.res_w1[r], p.res_b1[r] = C.ptr(blk.layer1.weight), C.ptr(blk.layer1.bias)
            p.res_g1[r], p.res_be1[r] = C.ptr(blk.bn1.weight), C.ptr(blk.bn1.bias)
            p.res_rm1[r], p.res_rv1[r] = C.ptr(blk.bn1.running_mean), C.ptr(blk.bn1.running_var)
            p.res_nbt1[r] = C.ptr(blk.bn1.num_batches_tracked)
            p.res_w2[r], p.res_b2[r] = C.ptr(blk.layer2.weight), C.ptr(blk.layer2.bias)
            p.res_g2[r], p.res_be2[r] = C.ptr(blk.bn2.weight), C.ptr(blk.bn2.bias)
            p.res_rm2[r], p.res_rv2[r] = C.ptr(blk.bn2.running_mean), C.ptr(blk.bn2.running_var)
            p.res_nbt2[r] = C.ptr(blk.bn2.num_batches_tracked)
        for l, cl in enumerate(self.cross_network):
            p.cross_w[l], p.cross_b[l] = C.ptr(cl.w.weight), C.ptr(cl.b)
        p.wf, p.bf = C.ptr(self.final_linear.weight), C.ptr(self.final_linear.bias)
        return p

    def _grad_struct(self, grads) -> C.Grads:
        g = C.Grads()
        it = iter(grads)
        g.user_table, g.item_table = C.ptr(next(it)), C.ptr(next(it))
        for i in range(len(self.cat_embeddings)):
            g.cat_table[i] = C.ptr(next(it))
        g.w0, g.b0 = C.ptr(next(it)), C.ptr(next(it))
        for r in range(len(self.res_blocks)):
            g.res_w1[r], g.res_b1[r], g.res_g1[r], g.res_be1[r] = (C.ptr(next(it)) for _ in range(4))
            g.res_w2[r], g.res_b2[r], g.res_g2[r], g.res_be2[r] = (C.ptr(next(it)) for _ in range(4))
        for l in range(len(self.cross_network)):
            g.cross_w[l], g.cross_b[l] = C.ptr(next(it)), C.ptr(next(it))
        g.wf, g.bf = C.ptr(next(it)), C.ptr(next(it))
        return g

    def _prep_inputs(self, user_ids, item_ids, cat_features, num_features):
        C.require_cuda(user_ids, item_ids, cat_features, num_features)
        n_cat = len(self.cat_embeddings)
        B = user_ids.numel()
        if item_ids.numel() != B or cat_features.shape[0] != B or num_features.shape[0] != B:
            raise RuntimeError("Sizes of tensors must match except in dimension 1")     # torch.cat's error
        if cat_features.dim() != 2 or cat_features.shape[1] < n_cat:
            raise IndexError("index out of range: cat_features has too few columns")
        if num_features.shape[1] != self._shape["n_num"]:
            raise RuntimeError("mat1 and mat2 shapes cannot be multiplied")
        user_ids = user_ids.reshape(-1).to(torch.int64).contiguous()
        item_ids = item_ids.reshape(-1).to(torch.int64).contiguous()
        cat_features = cat_features[:, :n_cat].to(torch.int64).contiguous()
        num_features = num_features.to(torch.float32).contiguous()
        return user_ids, item_ids, cat_features, num_features

    def parameters_to_allreduce(self):
        """Parameters whose ``.grad`` still has to be summed over the data-parallel ranks after backward(): all of
        them, except the user / item tables when their gradients were already built from the global batch."""
        comm = getattr(self, "_comm", None)
        sparse = comm is not None and comm.world > 1 and getattr(self, "dp_sparse_embedding_grads", True)
        skip = {id(self.user_embedding.weight), id(self.item_embedding.weight)} if sparse else set()
        return [p for p in self.parameters() if id(p) not in skip]

    def forward_rows(self, user_rows: torch.Tensor, item_rows: torch.Tensor, cat_features: torch.Tensor,
                     num_features: torch.Tensor) -> torch.Tensor:
        """forward() with the user / item embedding rows already fetched ([B, emb_dim] each, batch order) -- the
        entry used by ``distributed.RowShardedDCN`` after the all-to-all exchange.  The rows act as per-sample
        tables (id = batch position); gradients flow back into them."""
        C.require_cuda(user_rows, item_rows)
        B = user_rows.shape[0]
        ar = torch.arange(B, dtype=torch.int64, device=user_rows.device)
        self._row_override = (user_rows.contiguous(), item_rows.contiguous())
        try:
            return self.forward(ar, ar, cat_features, num_features)
        finally:
            self._row_override = None

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, user_ids: torch.Tensor, item_ids: torch.Tensor, cat_features: torch.Tensor,
                num_features: torch.Tensor) -> torch.Tensor:
        user_ids, item_ids, cat_features, num_features = self._prep_inputs(user_ids, item_ids, cat_features,
                                                                           num_features)
        B = user_ids.numel()
        dims = self._dims()
        batch = C.Batch(C.ptr(user_ids), C.ptr(item_ids), C.ptr(cat_features), C.ptr(num_features), B)
        if self.check_ids:
            flag = torch.zeros(1, dtype=torch.int32, device=user_ids.device)
            C.check(C.lib().dcnr_check_ids(dims, batch, C.ptr(flag), C.stream()))
        if self.training:
            self._flags(user_ids.device)
            if B == 1 and len(self.res_blocks) > 0 and dims.comm is None:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                                 f"torch.Size([1, {self._shape['hidden']}])")
            params = self._ordered_params()
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                logits = _DCNTrainFn.apply(self, user_ids, item_ids, cat_features, num_features, *params)
            else:
                with torch.no_grad():
                    logits = _DCNTrainFn.forward(_NullCtx(), self, user_ids, item_ids, cat_features, num_features,
                                                 *params)
        else:
            # eval(): folded BatchNorm, no dropout, nothing saved.  The result carries no autograd
            # graph (the reference only calls eval() forwards under torch.no_grad()).
            logits = self._eval_call(dims, batch, B, user_ids.device)
            if not self.defer_eval_checks and self.check_eval_flags():
                # an activation left the fp16 range of the fused tower: same batch on the tf32x3 kernels (still the GPU)
                dims.precision = C.PRECISIONS["tf32x3"]
                logits = self._eval_call(dims, batch, B, user_ids.device)
                if self.check_eval_flags():
                    raise RuntimeError("unexpected range flag from the tf32x3 path")
        return logits.squeeze()       # [B]; 0-d when B == 1, like train.py:170

    def _flags(self, dev):
        if self._eval_flags is None or self._eval_flags.device != dev:
            self._eval_flags = torch.zeros(4, dtype=torch.int32, device=dev)      # [0] flag bits, [1..2] diagnostics
        return self._eval_flags

    def _eval_call(self, dims, batch, B, dev):
        dims.eval_flags = C.ptr(self._flags(dev))
        pstruct = self._param_struct()
        dims.tower_pack = self._tower_pack_ptr(dims, pstruct, dev)
        ws = torch.empty(C.lib().dcnr_workspace_bytes(dims, B, 0), dtype=torch.uint8, device=dev)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        C.check(C.lib().dcnr_forward_eval(dims, pstruct, batch, C.ptr(logits), C.ptr(ws), ws.numel(), C.stream()))
        return logits

    def _tower_pack_ptr(self, dims, pstruct, dev):
        """The fused tower's prepared weights (dcnr_tower_prepare), rebuilt only when a tensor it is made from changed: torch
        bumps ``Tensor._version`` on every in-place update (optimizer step, load_state_dict, BatchNorm running statistics) and
        ``.to()`` / ``.cuda()`` move the storage, so (data_ptr, _version) of those tensors is the cache key."""
        if C.PRECISION_NAMES[dims.precision] not in ("fp16x3", "bf16") or not C.lib().dcnr_tower_eval_supported(dims):
            return None
        src = [self.initial_deep_layer.weight, self.initial_deep_layer.bias, self.final_linear.weight]
        for blk in self.res_blocks:
            for lin, bn in ((blk.layer1, blk.bn1), (blk.layer2, blk.bn2)):
                src += [lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = (dims.precision, float(dims.bn_eps), tuple((t.data_ptr(), t._version) for t in src))
        if self._tower_pack is None or self._tower_pack[0] != key:
            pack = torch.empty(C.lib().dcnr_tower_pack_bytes(dims), dtype=torch.uint8, device=dev)
            C.check(C.lib().dcnr_tower_prepare(dims, pstruct, dims.precision, C.ptr(pack), pack.numel(), C.stream()))
            self._tower_pack = (key, pack)
        return C.ptr(self._tower_pack[1])

    def check_eval_flags(self) -> bool:
        """Read (and clear) the flag word of the eval() forwards since the last check.  Raises ``IndexError`` like
        ``nn.Embedding`` if an id was out of range; returns True if a batch has to be re-run on the tf32x3 path because an
        activation left the fp16 range of the fused tower.  Synchronises the current stream."""
        if self._eval_flags is None:
            return False
        f = int(self._eval_flags[0].item())
        if f:
            self._eval_flags.zero_()
        if f & 1:
            raise IndexError("index out of range in self")
        return bool(f & 2)


class _NullCtx:
    """Stand-in autograd context for the no-grad train() forward."""
    needs_input_grad = ()
