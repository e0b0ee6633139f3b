"""Layer calls on HOST buffers, double buffered: while step i computes, step i+1's features travel host -> device and
step i-1's results travel device -> host (three CUDA streams, PCIe is full duplex).

The reference layer itself stages through the host on every call (GLL.py:27,30,73,90,133-137: X to numpy, Pred back,
grad_output to numpy, COO back).  Callers that hold their data on the host -- the evaluation path works on numpy arrays
(utils.py:570-593) -- pay one H2D of X and one D2H of (pred, dX) per call here; this class hides those copies behind the
kernels of the neighbouring calls.  Calls must be independent of each other (evaluation batches, attack restarts): a
training loop whose next features depend on this step's gradient keeps its tensors on the device and needs none of this.

    pipe = HostPipeline(n, d, k_lab, l, device, loss_fn=lambda pred, slot: ..., tau=0.0, epsilon="auto")
    for X_host, Y_host in batches:          # pinned host tensors
        pipe.submit(X_host, Y_host)
        if pipe.outstanding == pipe.depth:
            pred_host, dX_host = pipe.collect()
    while pipe.outstanding: pipe.collect()
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .GLL import LaplaceLearningSparseHard


class HostPipeline:
    def __init__(self, n: int, d: int, k_lab: int, l: int, device, loss_fn: Callable, tau: float = 0.0, epsilon="auto",
                 depth: int = 2, label_dtype: torch.dtype = torch.float32, layer: Optional[Callable] = None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device (graphlearninglayer_b200 has no CPU path)")
        self.depth, self.tau, self.epsilon, self.loss_fn = depth, tau, epsilon, loss_fn
        self.layer = layer or LaplaceLearningSparseHard.apply
        m = n - k_lab
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.Xd = [torch.empty((n, d), dtype=torch.float32, device=self.device).requires_grad_(True) for _ in range(depth)]
        self.Yd = [torch.empty((k_lab, l), dtype=label_dtype, device=self.device) for _ in range(depth)]
        self.pred_h = [torch.empty((m, l), dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.dX_h = [torch.empty((n, d), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self._keep = [None] * depth      # device results of a slot stay referenced until its D2H has been waited for
        self._used = [False] * depth
        self._head = 0                   # next slot to submit into
        self._tail = 0                   # oldest outstanding slot
        self.outstanding = 0
        self.h2d_bytes = n * d * 4 + k_lab * l * self.Yd[0].element_size()
        self.d2h_bytes = m * l * 8 + n * d * 4

    def submit(self, X_host: torch.Tensor, Y_host: torch.Tensor) -> int:
        """Enqueue one forward+backward on host inputs (pinned memory for asynchronous copies); returns the slot."""
        if self.outstanding == self.depth:
            raise RuntimeError("pipeline full: collect() before submitting more")
        s = self._head
        with torch.cuda.stream(self.s_in):
            if self._used[s]:
                self.s_in.wait_event(self.ev_comp[s])   # the previous call in this slot no longer reads Xd / Yd
            with torch.no_grad():
                self.Xd[s].copy_(X_host, non_blocking=True)
                self.Yd[s].copy_(Y_host, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.ev_in[s])
            if self._used[s]:
                self.s_comp.wait_event(self.ev_out[s])  # the slot's previous gradient has left the device
            self.Xd[s].grad = None
            pred = self.layer(self.Xd[s], self.Yd[s], self.tau, self.epsilon)
            loss = self.loss_fn(pred, s)
            loss.backward()
            self.ev_comp[s].record(self.s_comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[s])
            self.pred_h[s].copy_(pred.detach(), non_blocking=True)
            self.dX_h[s].copy_(self.Xd[s].grad, non_blocking=True)
            self.ev_out[s].record(self.s_out)
        self._keep[s] = (pred, loss)
        self._used[s] = True
        self._head = (s + 1) % self.depth
        self.outstanding += 1
        return s

    def collect(self):
        """Wait for the oldest outstanding call; returns (pred_host float64 m x l, dX_host float32 n x d) -- views of the
        slot's pinned buffers, valid until the slot is submitted into again."""
        if self.outstanding == 0:
            raise RuntimeError("nothing to collect")
        s = self._tail
        self.ev_out[s].synchronize()
        self._keep[s] = None
        self._tail = (s + 1) % self.depth
        self.outstanding -= 1
        return self.pred_h[s], self.dX_h[s]

    def drain(self):
        out = []
        while self.outstanding:
            out.append(self.collect())
        return out
