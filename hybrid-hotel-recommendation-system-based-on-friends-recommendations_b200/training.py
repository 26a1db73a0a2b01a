"""The training step (train.py:218-226: forward, BCEWithLogitsLoss, backward) as ONE CUDA graph.

At data-parallel batch sizes (8 192 - 32 768 rows per GPU) the ~100 kernels of a step are shorter than the host
time to launch them from Python, so the step is captured once -- forward, loss, backward and the NCCL gradient
exchange on the library's communicator -- and replayed.  Inputs live in static device tensors; dropout draws a
fresh mask on every replay from a device-side step counter (``dcnr_dims.dropout_step``); gradients land in the
parameters' ``.grad`` tensors (static across replays), so a stock ``torch.optim`` optimizer steps on them as usual.
"""
from __future__ import annotations

from typing import Optional

import torch

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
        return self.loss
