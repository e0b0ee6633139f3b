"""CPU numerics experiment for the streaming CG (C5, one GPU): would a bf16 copy of the iterate for the off-diagonal gathers pay?

The streaming kernel is bound by the gathered u rows that cross L2 -> SM; gathering bf16 instead of
fp32 would halve that volume.  This script restates the kernel's Jacobi-preconditioned multi-RHS CG in numpy (fp32 vectors,
fp64 dot products, per-column stop at |r|^2 <= tol^2 max|b|^2) and replaces the off-diagonal part of A u by A_off bf16(u),
optionally with the residual recomputed from an exact fp32 product every `rr` iterations (one extra full-precision SpMV each
time).  It reports iterations until the 1e-5 parity bar against the fp64 direct solve is met.

    python tools/cg_bf16_experiment.py            # C2-like and C4 systems (seconds)
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gll_oracle as O  # noqa: E402  (test infrastructure; this tool is a numerics study, not product code)


def bf16(a):
    """round-to-nearest-even bfloat16 of an fp32 array, returned as fp32"""
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((b >> 16) & 1) + np.uint32(0x7FFF)
    return ((b + r) & np.uint32(0xFFFF0000)).view(np.float32)


def jacobi_cg(Luu, B, mode, rr, tol=1e-7, max_iter=120, ref=None, bar=1e-5):
    """mode 'fp32': the kernel's arithmetic; 'bf16': off-diagonal gathers of bf16(u).  Returns (iterations run, iteration at
    which the parity bar was first met or None, final max relative error)."""
    A = Luu.tocsr().astype(np.float32)
    dg = A.diagonal().astype(np.float32)
    Aoff = (A - sp.diags(dg)).tocsr().astype(np.float32)
    dinv = (1.0 / dg).astype(np.float32)
    b = B.astype(np.float32)
    x = np.zeros_like(b)
    r = b.copy()
    p = np.zeros_like(b)
    s = np.zeros_like(b)
    tol2 = tol * tol * float((b.astype(np.float64) ** 2).sum(axis=0).max())
    g_old = np.ones(b.shape[1])
    a_old = np.ones(b.shape[1])
    met = None
    frozen = np.zeros(b.shape[1], dtype=bool)
    refn = np.abs(ref).max()

    def matvec(u, exact):
        uo = u if (exact or mode == "fp32") else bf16(u)
        return (dg[:, None] * u + Aoff @ uo).astype(np.float32)

    for it in range(max_iter):
        if rr and it > 0 and it % rr == 0:  # residual replacement: r = b - A x with the exact product
            r = (b - (dg[:, None] * x + Aoff @ x)).astype(np.float32)
        u = (r * dinv[:, None]).astype(np.float32)
        w = matvec(u, False)
        g = (r.astype(np.float64) * u).sum(axis=0)
        d = (w.astype(np.float64) * u).sum(axis=0)
        rr2 = (r.astype(np.float64) ** 2).sum(axis=0)
        err = np.abs(x.astype(np.float64) - ref).max() / refn
        if met is None and err <= bar:
            met = it
        live = (rr2 > tol2) & ~frozen  # per-column stop test and freeze, as in the kernel
        if not live.any():
            return it, met, err
        beta = np.where(it == 0, 0.0, g / g_old)
        den = d - beta * g / a_old
        ok = live & (den > 0) & (g > 0)
        frozen |= live & ~ok  # breakdown at the rounding floor: the column stops moving
        alpha = np.where(ok, g / np.where(ok, den, 1.0), 0.0)
        beta = np.where(ok, beta, 0.0)
        g_old, a_old = np.where(ok, g, g_old), np.where(ok, alpha, a_old)
        p = (u + beta.astype(np.float32) * p).astype(np.float32)
        s = (w + beta.astype(np.float32) * s).astype(np.float32)
        x = (x + alpha.astype(np.float32) * p).astype(np.float32)
        r = (r - alpha.astype(np.float32) * s).astype(np.float32)
    err = np.abs(x.astype(np.float64) - ref).max() / refn
    return max_iter, met, err


def system(seed, k_lab, m, d, l, sigma):
    X, Y, _, _ = O.synth_inputs(seed, k_lab, m, d, l, sigma)
    g = O.build_graph(X, 25, "auto")
    Luu, B = O.laplace_system(g.W, Y, 0.0)[:2]
    return Luu.tocsr(), np.asarray(B)


if __name__ == "__main__":
    cases = [("c4-like (2048 + 6144, d = 128)", (2, 2048, 6144, 128, 10, 4.5)),
             ("dense unlabeled block (512 + 7680, d = 64, 20 classes)", (5, 512, 7680, 64, 20, 3.0))]
    for name, args in cases:
        Luu, B = system(*args)
        ref = spla.splu(Luu.tocsc().astype(np.float64)).solve(B.astype(np.float64))
        print(f"== {name}: m = {Luu.shape[0]}, nnz = {Luu.nnz}")
        for mode, rr in (("fp32", 0), ("bf16", 0), ("bf16", 8), ("bf16", 4), ("bf16", 2)):
            it, met, err = jacobi_cg(Luu, B, mode, rr, ref=ref)
            print(f"   gathers {mode:5s} residual replacement every {rr or '-':>2}: stop after {it:3d} iterations, parity 1e-5 first met at "
                  f"{met if met is not None else 'never'}, final max rel error {err:.2e}")
