"""The training step (train.py:218-226: forward, BCEWithLogitsLoss, backward) as ONE CUDA graph.

At data-parallel batch sizes (8 192 - 32 768 rows per GPU) the ~100 kernels of a step are shorter than the host
time to launch them from Python, so the step is captured once -- forward, loss, backward and the NCCL gradient
exchange on the library's communicator -- and replayed.  Inputs live in static device tensors; dropout draws a
fresh mask on every replay from a device-side step counter (``dcnr_dims.dropout_step``); gradients land in the
parameters' ``.grad`` tensors (static across replays, re-attached after every replay), so a stock ``torch.optim``
optimizer loop -- ``zero_grad()`` with its default ``set_to_none=True`` included -- steps on them as usual.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi as C
from . import functional as F_
from .distributed import Communicator, allreduce_gradients, attach


class GraphedTrainStep:
    def __init__(self, model, batch_size: int, comm: Optional[Communicator] = None, warmup: int = 3):
        self.model, self.comm = model, comm
        self.world = comm.world if comm is not None else 1
        dev = next(model.parameters()).device
        s = model._shape
        B = int(batch_size)
        self.user_ids = torch.zeros(B, dtype=torch.int64, device=dev)
        self.item_ids = torch.zeros(B, dtype=torch.int64, device=dev)
        self.cat = torch.zeros((B, len(s["cat_rows"])), dtype=torch.int64, device=dev)
        self.num = torch.zeros((B, s["n_num"]), dtype=torch.float32, device=dev)
        self.labels = torch.zeros(B, dtype=torch.float32, device=dev)
        model.train()
        attach(model, comm)
        model._dropout_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.params = list(model.parameters())
        self.reduce_params = model.parameters_to_allreduce()
        self.graph = None
        self.loss = None
        self._warmup = warmup

    def _step(self):
        logits = self.model(self.user_ids, self.item_ids, self.cat, self.num)
        loss, dl = F_.bce_with_logits(logits.detach(), self.labels)
        if self.world > 1:
            dl = dl / self.world                              # mean over the GLOBAL batch
        logits.backward(gradient=dl)
        allreduce_gradients(self.reduce_params, comm=self.comm, average=False)
        return loss

    def capture(self):
        """Warm up on a side stream, then capture one step.  The warm-up steps are REAL train-mode steps on whatever the
        static inputs hold, so everything they mutate besides the gradients -- BatchNorm running statistics,
        ``num_batches_tracked``, the device dropout counter -- is snapshotted first and restored afterwards: capturing
        leaves the model exactly as it found it (the reference's loop updates those once per batch, train.py:218-226)."""
        snap = {n: b.detach().clone() for n, b in self.model.named_buffers()}
        step0 = self.model._dropout_step.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                for p in self.params:
                    p.grad = None
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        # the gradient tensors the graph writes on every replay; re-attached after each replay in case the training loop
        # detached them (optimizer.zero_grad() defaults to set_to_none=True, train.py:222)
        self._grads = [p.grad for p in self.params]
        with torch.no_grad():
            for n, b in self.model.named_buffers():
                b.copy_(snap[n])
            self.model._dropout_step.copy_(step0)
        return self

    def load(self, user_ids, item_ids, cat_features, num_features, labels):
        self.user_ids.copy_(user_ids, non_blocking=True)
        self.item_ids.copy_(item_ids, non_blocking=True)
        self.cat.copy_(cat_features, non_blocking=True)
        self.num.copy_(num_features, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)

    def __call__(self, user_ids=None, item_ids=None, cat_features=None, num_features=None, labels=None):
        """One training step on the given batch (or on whatever ``load`` put in the static inputs); returns the
        loss tensor (device scalar, overwritten by the next step)."""
        if user_ids is not None:
            self.load(user_ids, item_ids, cat_features, num_features, labels)
        if self.graph is None:
            self.capture()
        self.graph.replay()
        C.mark_mutated(self.model.buffers())                 # the replay updated the BatchNorm statistics behind torch's back
        for p, g in zip(self.params, self._grads):
            if p.grad is not g:
                p.grad = g
        return self.loss
