"""`custom_ce_loss(softmax_logits, targets)` of the reference (losses.py:128-136; pasted again at
train_and_adversarial.py:458 and adversarial.py:453) as ONE kernel: the loss and its gradient with respect to the
layer's output come out of the same launch (the reference spends a dozen small PyTorch kernels on one_hot / add / log / mul /
sum / neg / div and their backward).  Same name, arguments and result as the reference function; CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib

__all__ = ["custom_ce_loss", "LaplaceLearningCELoss", "laplace_ce_loss"]


class _CELoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, softmax_logits, targets):
        if not softmax_logits.is_cuda:
            raise RuntimeError("graphlearninglayer_b200.losses.custom_ce_loss needs CUDA tensors (no CPU path)")
        p = softmax_logits.detach()
        if p.dtype not in (torch.float32, torch.float64):
            p = p.float()
        p = p.contiguous()
        m, l = p.shape
        t = targets.detach().to(device=p.device, dtype=torch.int64).contiguous()
        if t.numel() != m:
            raise ValueError("targets must hold one class index per row of softmax_logits")
        loss = torch.empty((), dtype=p.dtype, device=p.device)
        grad = torch.empty_like(p)
        wsb = lib.gll_ce_loss_workspace_bytes(m)
        ws = torch.empty(wsb, dtype=torch.uint8, device=p.device) if wsb else None
        with torch.cuda.device(p.device):
            _lib.check(lib.gll_ce_loss(p.data_ptr(), int(p.dtype == torch.float64), t.data_ptr(), m, l, loss.data_ptr(),
                                       grad.data_ptr(), None, ws.data_ptr() if wsb else None, wsb,
                                       torch.cuda.current_stream(p.device).cuda_stream), "gll_ce_loss")
        ctx.save_for_backward(grad)
        ctx.in_dtype = softmax_logits.dtype
        return loss.to(softmax_logits.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        return (grad * grad_output.to(grad.dtype)).to(ctx.in_dtype), None


def custom_ce_loss(softmax_logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """-sum(one_hot(targets) * log(softmax_logits + 1e-8)) / batch_size  (losses.py:128-136)."""
    return _CELoss.apply(softmax_logits, targets)


class LaplaceLearningCELoss(torch.autograd.Function):
    """The layer with the loss behind it in ONE autograd node:

        pred = LaplaceLearningSparseHard.apply(features, label_matrix, tau, epsilon)      # FullySup.py:156
        loss = custom_ce_loss(pred, targets)                                               # FullySup.py:158, losses.py:128-136

    becomes ``loss, pred = LaplaceLearningCELoss.apply(features, label_matrix, targets, tau, epsilon)``.  The loss kernel leaves
    d loss / d pred on the device; backward hands it to the adjoint solve as its right-hand side together with the incoming
    (scalar) gradient, which the solver kernel multiplies in itself -- no elementwise kernels between the loss and the solve, no
    host read.  Same numbers as the two calls.  ``pred`` is returned for accuracy bookkeeping and carries no gradient."""

    @staticmethod
    def forward(ctx, X, label_matrix, targets, tau=0, epsilon="auto"):
        from .GLL import _forward_impl, _stream_ptr

        pred, st, Xc = _forward_impl(X, label_matrix, tau, epsilon)
        m, l = pred.shape
        t = targets.detach().to(device=pred.device, dtype=torch.int64).contiguous()
        if t.numel() != m:
            raise ValueError("targets must hold one class index per unlabeled row")
        loss = torch.empty((), dtype=pred.dtype, device=pred.device)
        grad = torch.empty_like(pred)
        wsb = lib.gll_ce_loss_workspace_bytes(m)
        ws = torch.empty(wsb, dtype=torch.uint8, device=pred.device) if wsb else None
        with torch.cuda.device(pred.device):
            _lib.check(lib.gll_ce_loss(pred.data_ptr(), int(pred.dtype == torch.float64), t.data_ptr(), m, l, loss.data_ptr(),
                                       grad.data_ptr(), None, ws.data_ptr() if wsb else None, wsb, _stream_ptr(pred.device)),
                       "gll_ce_loss")
        ctx.gll_state = st
        ctx.x_dtype = X.dtype
        ctx.save_for_backward(Xc, grad)
        ctx.mark_non_differentiable(pred)
        ctx.set_materialize_grads(True)
        return loss, pred

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_pred):
        from .GLL import _backward_impl

        Xc, grad = ctx.saved_tensors
        dX = _backward_impl(ctx.gll_state, Xc, grad, scale=grad_loss)
        if dX.dtype != ctx.x_dtype:
            dX = dX.to(ctx.x_dtype)
        return dX, None, None, None, None


def laplace_ce_loss(features, label_matrix, targets, tau=0, epsilon="auto"):
    """``custom_ce_loss(LaplaceLearningSparseHard.apply(features, label_matrix, tau, epsilon), targets)`` as one node: (loss, pred)."""
    return LaplaceLearningCELoss.apply(features, label_matrix, targets, tau, epsilon)
