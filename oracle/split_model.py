"""numpy model of the operand split behind the tensor-core kNN candidate search -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The product kernel (graphlearninglayer_b200/csrc/knn.cu: sqnorm_split_f16_kernel; knn_tc.cu: knn_gram_topk_tc_kernel,
knn_tc_err_coef; knn.cu: knn_err_bound) replaces the distances behind ``gl.weightmatrix.knnsearch`` (GLL.py:181-189) by
approximate ones that only SELECT candidates; exactness of the emitted lists rests on a proven bound
``|d~^2_ij - d^2_ij| <= coef (|x_i|^2 + max|x|^2) + 2 R`` with ``R = rho (|x_i| + max|x| + rho)`` for the default one-pass
Gram ``hi_i . hi_j`` and ``R = |x_i| rho`` for the two-pass ``(hi_i + lo_i) . hi_j``.  This file restates the split and the bound in numpy so
that tests/test_split_bound.py can check the bound itself on the CPU, for feature scales and shapes the GPU tests cannot
sweep.  The fp32 accumulation order inside the tensor core is not modelled (it has its own budget inside ``coef``); the
Gram entries here are the exact fp64 products of the split operands, and, as a second variant, numpy's fp32 matmul.
"""
from __future__ import annotations

import math

import numpy as np


F16_TARGET_LOG2 = 8  # rows are scaled to a norm in [0.58, 1.16) * 2^8 (knn.cu)


def scale_exponent(sq64: np.ndarray) -> np.ndarray:
    """E_i of sqnorm_split_f16_kernel: from the bits of float(1.5 |x_i|^2), minus the target exponent; zero / non-finite rows take
    the unit-norm bucket; clamped to +-60."""
    t = (1.5 * sq64).astype(np.float32)
    bits = t.view(np.uint32)
    ex = ((bits >> 23) & 0xFF).astype(np.int64)
    E = np.floor_divide(ex - 127 + 1, 2) - F16_TARGET_LOG2
    E = np.clip(E, -60, 60)
    E[(bits == 0) | (ex == 0xFF)] = -F16_TARGET_LOG2
    return E


def split_f16(X: np.ndarray):
    """Returns hi, lo (float16, scaled rows), E (int), sq (float32 |x_i|^2) and rho (float32, rounded up)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    s = (X.astype(np.float64) ** 2).sum(axis=1)
    sq = s.astype(np.float32)
    E = scale_exponent(s)
    down = np.ldexp(np.float32(1.0), -E).astype(np.float32)[:, None]
    z = X * down  # exact: powers of two
    hi = z.astype(np.float16)
    lo = (z - hi.astype(np.float32)).astype(np.float16)
    resid = X.astype(np.float64) - np.ldexp(hi.astype(np.float64), E[:, None])
    rho64 = math.sqrt(float((resid ** 2).sum(axis=1).max())) * 1.000001
    rho = np.float32(rho64)
    if float(rho) < rho64:
        rho = np.nextafter(rho, np.float32(np.inf))
    return hi, lo, E, sq, rho


def flush_fp16_subnormals(h: np.ndarray) -> np.ndarray:
    """What a multiplier WITHOUT fp16 subnormal support would see (the bound must hold for it too)."""
    out = h.copy()
    out[np.abs(out.astype(np.float32)) < np.float32(2.0 ** -14)] = 0
    return out


def approx_d2_f16(hi, lo, E, sq, passes: int = 1, fp32_accumulate: bool = False, flush_subnormals: bool = False) -> np.ndarray:
    """d~^2_ij = |x_i|^2 + (|x_j|^2 + (acc 2^E_i)(-2 2^E_j)) as the epilogue forms it, with acc = hi_i . hi_j (one pass) or
    (hi_i + lo_i) . hi_j (two passes)."""
    if flush_subnormals:
        hi, lo = flush_fp16_subnormals(hi), flush_fp16_subnormals(lo)
    if passes == 1:
        lo = np.zeros_like(lo)
    if fp32_accumulate:
        a = hi.astype(np.float32) @ hi.astype(np.float32).T + lo.astype(np.float32) @ hi.astype(np.float32).T
        acc = a.astype(np.float32)
    else:
        a = (hi.astype(np.float64) + lo.astype(np.float64)) @ hi.astype(np.float64).T
        acc = a.astype(np.float32)  # the accumulator is fp32
    ri = np.ldexp(np.float32(1.0), E).astype(np.float32)
    cj = (np.float32(-2.0) * ri).astype(np.float32)
    t = (acc * ri[:, None]).astype(np.float32)
    key = (t.astype(np.float64) * cj[None, :].astype(np.float64) + sq[None, :].astype(np.float64)).astype(np.float32)  # one fma
    return (key + sq[:, None]).astype(np.float32)


def epilogue_values_packed(hi, lo, E, sq, passes: int = 1, fp32_accumulate: bool = True):
    """What the tensor-core epilogue STORES and FLUSHES since round 2 (knn_tc.cu, packed keys): the column value is shifted by
    Cs = 2 max|x|^2 (so that it is positive and its bits order like unsigned integers), v' = fl(t cj + fl(|x_j|^2 + Cs)) with
    t = acc 2^E_i (exact), five mantissa bits are cleared, and a flushed candidate carries fl(fl(v'_cleared - Cs) + |x_i|^2).
    Returns (flushed (n, n) float32, cleared v' (n, n) float32): selection inside a row happens on the cleared values."""
    if passes == 1:
        lo = np.zeros_like(lo)
    if fp32_accumulate:
        acc = (hi.astype(np.float32) @ hi.astype(np.float32).T + lo.astype(np.float32) @ hi.astype(np.float32).T).astype(np.float32)
    else:
        acc = ((hi.astype(np.float64) + lo.astype(np.float64)) @ hi.astype(np.float64).T).astype(np.float32)
    ri = np.ldexp(np.float32(1.0), E).astype(np.float32)
    cj = (np.float32(-2.0) * ri).astype(np.float32)
    cs = np.float32(2.0) * np.float32(sq.max())
    sj = (sq + cs).astype(np.float32)                                            # staged column value, one rounding
    t = (acc * ri[:, None]).astype(np.float32)                                   # exact: power of two
    v = (t.astype(np.float64) * cj[None, :].astype(np.float64) + sj[None, :].astype(np.float64)).astype(np.float32)  # one FMA
    cleared = (v.view(np.uint32) & np.uint32(0xFFFFFFE0)).view(np.float32)
    flushed = ((cleared - cs).astype(np.float32) + sq[:, None]).astype(np.float32)
    return flushed, cleared


def err_coef(d: int, passes: int = 1) -> float:
    """knn_tc_err_coef (knn_tc.cu)."""
    steps = passes * math.ceil(d / 16) + 8.0
    split = (2.0 if passes == 1 else 1.0) * (1.0 + math.sqrt(d)) / 2097152.0 * 1.01
    e = split + steps * 4.76837158203125e-7 + 12.0 * 5.9604644775390625e-8
    return float(np.float32(4.0 * e))


def err_bound(d: int, sq: np.ndarray, rho, passes: int = 1) -> np.ndarray:
    """knn_err_bound (knn.cu), per row i."""
    sq64 = sq.astype(np.float64)
    xi, xm, r = np.sqrt(sq64), math.sqrt(float(sq.max())), float(rho)
    rterm = r * (xi + xm + r) if passes == 1 else xi * r
    return err_coef(d, passes) * (sq64 + float(sq.max())) + 2.0 * rterm * 1.000001


def exact_d2(X: np.ndarray) -> np.ndarray:
    Xd = X.astype(np.float64)
    g = Xd @ Xd.T
    s = (Xd ** 2).sum(axis=1)
    return s[:, None] + s[None, :] - 2.0 * g  # fp64 Gram form: its own error (1e-16 relative to the norms) is far below the bounds tested


# ---------------------------------------------------------------------------------------------------------------------------
# the selection pipeline on top of the approximate distances (knn_rerank_kernel / knn_fallback_kernel in knn.cu)
# ---------------------------------------------------------------------------------------------------------------------------
def select_rerank_prove(X: np.ndarray, approx: np.ndarray, bound: np.ndarray, k: int = 25, kc: int = 32):
    """Restates what the GPU does with the approximate distances: per row the kc candidates with the smallest (approximate
    distance, index) key (self excluded), exact fp64 re-rank of those candidates by (distance, index), and the completeness
    proof ``approx(kc-th candidate) - bound > exact (k-1)-th distance``.  Returns (ind (n,k) with self in slot 0,
    proven (n,) bool).  Rows that are not proven go to the brute-force fallback on the GPU, so only proven rows have to
    be right here."""
    Xd = np.ascontiguousarray(X, dtype=np.float32).astype(np.float64)
    n = Xd.shape[0]
    a = approx.astype(np.float32).copy()
    a[np.arange(n), np.arange(n)] = np.inf
    ind = np.empty((n, k), dtype=np.int64)
    proven = np.zeros(n, dtype=bool)
    for i in range(n):
        order = np.lexsort((np.arange(n), a[i]))[:kc]
        order = order[np.isfinite(a[i][order])]
        lower = np.float32(np.inf) if len(order) < kc else a[i][order[-1]]
        d2 = ((Xd[i][None, :] - Xd[order]) ** 2).sum(axis=1)
        rr = np.lexsort((order, d2))[: k - 1]
        ind[i, 0] = i
        ind[i, 1:1 + len(rr)] = order[rr]
        dk = d2[rr[-1]]
        proven[i] = bool(np.isinf(lower) or (float(lower) - float(bound[i]) > dk))
    return ind, proven
