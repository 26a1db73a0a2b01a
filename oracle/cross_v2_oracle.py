"""CPU statement of the OPT-IN DCN-v2 ("full-matrix") cross network.  TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path.

PARITY UNPINNED against the reference: the reference's CrossLayer (train.py:90-99 / main.py:61-70) is the rank-1 form
x (1 + x.w) + b, restated in oracle/dcnr_oracle.py and pinned to the reference's own classes.  The full-matrix form
    x_{l+1} = x0 * (W_l x_l + b_l) + x_l
is what BASELINE.json's north_star sentence (and Documentation.md:100's "x0 * (w^T x_l) + b + x_l" prose) describes; it has
different parameters (W_l is [D, D]), so the only possible anchor is this plain statement of the published DCN-v2 formula
(Wang et al. 2021, eq. 1), written two independent ways: torch autograd in float64 and a numpy closed form for the
gradients.  tests/test_oracle.py checks the two against each other.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def cross_v2_forward_torch(x0: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> torch.Tensor:
    """x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l with nn.Linear-layout weights ([out, in])."""
    x = x0
    for w, b in zip(weights, biases):
        x = x0 * (x @ w.t() + b) + x
    return x


def cross_v2_numpy(x0: np.ndarray, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray], gy: np.ndarray
                   ) -> Tuple[np.ndarray, np.ndarray, List[np.ndarray], List[np.ndarray]]:
    """float64 closed form: (y, dL/dx0, [dL/dW_l], [dL/db_l]) for upstream gy -- the kernels' algorithm line by line
    (gm = g * x0 ; dx0 += g * u ; dx = gm W + g ; dW = gm^T x ; db = sum gm)."""
    x0 = x0.astype(np.float64)
    ws = [w.astype(np.float64) for w in weights]
    bs = [b.astype(np.float64) for b in biases]
    xs, us = [x0], []
    for w, b in zip(ws, bs):
        u = xs[-1] @ w.T + b
        us.append(u)
        xs.append(x0 * u + xs[-1])
    g = gy.astype(np.float64)
    dx0 = np.zeros_like(x0)
    gws, gbs = [None] * len(ws), [None] * len(ws)
    for l in reversed(range(len(ws))):
        gm = g * x0
        dx0 += g * us[l]
        gws[l] = gm.T @ xs[l]
        gbs[l] = gm.sum(0)
        g = gm @ ws[l] + g
    dx0 += g
    return xs[-1], dx0, gws, gbs
