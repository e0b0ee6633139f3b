"""CUDA-graph replay of the layer for static shapes.

    step = GraphedStep(n, d, k_lab, l, device, tau=0.0, epsilon="auto", loss="ce")
    loss = step(X, Y, y_query)        # forward + loss + backward as TWO graph launches; step.dX holds dL/dX afterwards
    pred = step.pred                  # (m, l) float64, the layer's output of the last call
    step = GraphedStep(..., loss_head=True)   # the same with layer + custom_ce_loss fused into one node (losses.laplace_ce_loss)

Every kernel of ``gll_forward`` / ``gll_backward`` is enqueued on the caller's stream with host-known launch
configurations, worst-case allocations and no host synchronisation (include/gll_b200.h), so the whole call can be
captured once and replayed: ~14 launches (two of them cooperative) collapse into two graph launches, which removes the
host side of a step (0.3-0.4 ms of Python, ctypes and launch calls -- what bounds several processes sharing one host).
Capture goes through ``torch.cuda.make_graphed_callables``; buffers the layer allocates while capturing live in the
graph's private pool.  Inputs are copied into static tensors, so shapes, dtypes, ``tau`` and ``epsilon`` are fixed per
instance; use one instance per shape.  The reference has no counterpart (its path synchronises with the host several times
per call, GLL.py:27,30,73,90); callers that keep the reference's call site use ``LaplaceLearningSparseHard.apply``.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .GLL import LaplaceLearningSparseHard
from .losses import custom_ce_loss


class GraphedStep:
    def __init__(self, n: int, d: int, k_lab: int, l: int, device, tau: float = 0.0, epsilon="auto",
                 loss: Optional[Callable] = None, label_dtype: torch.dtype = torch.float32, layer: Optional[Callable] = None,
                 warmup_inputs=None, loss_head: bool = False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA device (graphlearninglayer_b200 has no CPU path)")
        m = n - k_lab
        layer = layer or LaplaceLearningSparseHard.apply
        loss_fn = loss or custom_ce_loss  # custom_ce_loss(pred, target), losses.py:128-136

        def step(X, Y, yq):
            if loss_head:  # layer + custom_ce_loss as ONE autograd node (losses.LaplaceLearningCELoss)
                from .losses import laplace_ce_loss

                return laplace_ce_loss(X, Y, yq, tau, epsilon)
            pred = layer(X, Y, tau, epsilon)
            return loss_fn(pred, yq), pred

        # representative inputs for the warm-up / capture runs (the graph replays on whatever is copied in later)
        if warmup_inputs is None:
            g = torch.Generator(device="cpu").manual_seed(0)
            X0 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
            Y0 = torch.nn.functional.one_hot(torch.arange(k_lab) % l, l).to(label_dtype)
            y0 = torch.randint(0, l, (m,), generator=g)
        else:
            X0, Y0, y0 = warmup_inputs
        with torch.cuda.device(self.device):
            sample = (X0.to(self.device, torch.float32).requires_grad_(True), Y0.to(self.device, label_dtype),
                      y0.to(self.device, torch.int64))
            self._graphed = torch.cuda.make_graphed_callables(step, sample, num_warmup_iters=3)
        self.n, self.d, self.k_lab, self.l = n, d, k_lab, l
        self.dX = None
        self.pred = None

    def __call__(self, X: torch.Tensor, Y: torch.Tensor, y_query: torch.Tensor) -> torch.Tensor:
        """Forward + loss + backward.  Returns the (detached) loss; ``self.pred`` and ``self.dX`` hold the outputs (static
        tensors, overwritten by the next call)."""
        Xr = X if X.requires_grad else X.detach().requires_grad_(True)
        loss, pred = self._graphed(Xr, Y, y_query)
        (self.dX,) = torch.autograd.grad(loss, Xr)
        self.pred = pred.detach()
        return loss.detach()
