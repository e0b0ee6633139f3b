"""CPU oracle for the GLL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under
``graphlearninglayer_b200/`` imports it: the product path has no CPU fallback.

What it is: an fp64 numpy/scipy *restatement* of the algorithm in the reference's
``GLL.py`` (forward ``GLL.py:13-73``, backward ``GLL.py:75-177``, graph and weights
``GLL.py:180-244``, CG ``GLL.py:247-276``).  It never materialises the dense n x n
``C`` of ``GLL.py:209-213`` (it keeps the map kappa(i) instead), and it evaluates the
backward edge-by-edge instead of the per-class ``graph.gradient`` loop of
``GLL.py:111-120``.

Parity status: the reference ships no tests and no golden vectors, and its kNN
dependency (graphlearning -> annoy, approximate) is absent from this image, so parity
at the kNN boundary is **unpinned** by the reference itself.  The pin adopted here is:
unmodified ``/root/reference/GLL.py`` executed with the exact-kNN ``graphlearning``
stand-in in ``oracle/shim`` -> committed fixtures in ``tests/golden`` (generator:
``oracle/make_golden.py``).  ``tests/test_oracle.py`` checks this restatement against
those fixtures.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Union

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

K_DEFAULT = 25  # hard-coded at GLL.py:27


# --------------------------------------------------------------------------------------
# synthetic inputs: Gaussian clusters, L2-normalised, base rows first
# --------------------------------------------------------------------------------------
from graphlearninglayer_b200.synth import synth_inputs  # noqa: E402,F401  (numpy-only generator shared with bench.py)


def ce_loss_and_grad(pred: np.ndarray, y_query: np.ndarray):
    """custom_ce_loss of losses.py:128-136 and its gradient w.r.t. pred (fp64)."""
    mrows = pred.shape[0]
    p = pred[np.arange(mrows), y_query]
    loss = -np.sum(np.log(p + 1e-8)) / mrows
    g = np.zeros_like(pred, dtype=np.float64)
    g[np.arange(mrows), y_query] = -1.0 / (mrows * (p + 1e-8))
    return loss, g


# --------------------------------------------------------------------------------------
# exact kNN: the limit that annoy (GLL.py:181-183) approximates
# --------------------------------------------------------------------------------------
def exact_knn(X: np.ndarray, k: int = K_DEFAULT, block: int = 2048, slack: int = 8):
    """k nearest neighbours incl. self in slot 0 (distance 0), as knnsearch returns them.

    Candidates come from an fp64 Gram identity; the kept distances are recomputed as
    sqrt(sum((x_i - x_j)^2)) in fp64 from the fp32 inputs and rounded to fp32 (annoy stores
    and returns fp32).  Order: (distance, index) ascending.  Returns (ind int64, dist float64).
    """
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    Xd = X32.astype(np.float64)
    n = Xd.shape[0]
    if n < k:
        raise ValueError(f"need at least k={k} points, got {n}")
    sq = np.einsum("ij,ij->i", Xd, Xd)
    kc = min(n, k + slack)
    ind = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    for s in range(0, n, block):
        e = min(n, s + block)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (Xd[s:e] @ Xd.T)
        d2[np.arange(e - s), np.arange(s, e)] = -1.0  # self always first
        cand = np.argpartition(d2, kc - 1, axis=1)[:, :kc]
        diff = Xd[s:e, None, :] - Xd[cand]
        d2c = np.einsum("ijk,ijk->ij", diff, diff)
        rows = np.arange(s, e)[:, None]
        d2c[cand == rows] = -1.0
        order = np.lexsort((cand, d2c), axis=1)[:, :k]
        ci = np.take_along_axis(cand, order, axis=1)
        cd = np.take_along_axis(d2c, order, axis=1)
        cd[:, 0] = 0.0
        ind[s:e] = ci
        dist[s:e] = np.sqrt(np.maximum(cd, 0.0)).astype(np.float32).astype(np.float64)
    return ind, dist


def exact_knn_rows(X: np.ndarray, rows, k: int = K_DEFAULT, slack: int = 8):
    """exact_knn for a SAMPLE of rows against all points (full-size checks: n = 131072 and up, where the all-rows search
    would take minutes on the host).  Same candidates / exact recompute / (distance, index) order as exact_knn."""
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    rows = np.asarray(rows, dtype=np.int64)
    n = X32.shape[0]
    sq = np.einsum("ij,ij->i", X32, X32, dtype=np.float64)
    Xr = X32[rows].astype(np.float64)
    kc = min(n, k + slack)
    d2 = sq[rows, None] + sq[None, :]
    for s in range(0, n, 16384):  # the fp64 copy of X is only materialised block by block
        e = min(n, s + 16384)
        d2[:, s:e] -= 2.0 * (Xr @ X32[s:e].astype(np.float64).T)
    d2[np.arange(len(rows)), rows] = -1.0
    cand = np.argpartition(d2, kc - 1, axis=1)[:, :kc]
    diff = Xr[:, None, :] - X32[cand].astype(np.float64)
    d2c = np.einsum("ijk,ijk->ij", diff, diff)
    d2c[cand == rows[:, None]] = -1.0
    order = np.lexsort((cand, d2c), axis=1)[:, :k]
    ci = np.take_along_axis(cand, order, axis=1)
    cd = np.take_along_axis(d2c, order, axis=1)
    cd[:, 0] = 0.0
    return ci, np.sqrt(np.maximum(cd, 0.0)).astype(np.float32).astype(np.float64)


# --------------------------------------------------------------------------------------
# graph + weights (GLL.py:180-244)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Graph:
    n: int
    knn_ind: np.ndarray      # (n,k) int64
    knn_dist: np.ndarray     # (n,k) float64 (fp32-valued)
    dist: sp.csr_matrix      # symmetrised union graph, one distance per edge, sorted columns
    eps: np.ndarray          # (n,) float64
    kappa: Optional[np.ndarray]  # (n,) int64 for epsilon='auto' else None
    W: sp.csr_matrix
    V: sp.csr_matrix
    modV: Optional[sp.csr_matrix]


def build_graph(X: np.ndarray, k: int = K_DEFAULT, epsilon: Union[str, float] = "auto",
                knn: Optional[tuple] = None) -> Graph:
    """Union kNN graph, bandwidths and the three per-edge quantities.

    GLL.py:196-198: sparse Dist from the kNN lists, elementwise max with its transpose,
    exact zeros (self loops, duplicate points) are not edges.
    GLL.py:205: eps_i (auto) = distance from i to the last kNN entry kappa(i).
    GLL.py:216-218 / 233-234: W = exp(-4 d^2/(eps_i eps_j)), V = -8 W/(eps_i eps_j),
    modV = d^2 V / (2 eps_i^2).
    """
    knn_ind, knn_dist = exact_knn(X, k) if knn is None else knn
    n, kk = knn_ind.shape
    rows = np.repeat(np.arange(n), kk)
    D = sp.coo_matrix((knn_dist.ravel(), (rows, knn_ind.ravel())), shape=(n, n)).tocsr()
    D = D.maximum(D.T).tocsr()
    D.eliminate_zeros()
    D.sort_indices()
    r = np.repeat(np.arange(n), np.diff(D.indptr))
    c = D.indices
    v = D.data
    if isinstance(epsilon, str):
        if epsilon != "auto":
            raise ValueError("epsilon must be a float or 'auto'")
        kappa = knn_ind[:, -1].astype(np.int64)
        eps = np.asarray(D[knn_ind[:, 0], kappa]).ravel().astype(np.float64)
    else:
        kappa = None
        eps = float(epsilon) * np.ones(n)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        w = np.exp(-4.0 * v * v / eps[r] / eps[c])
        vv = -8.0 * w / eps[r] / eps[c]
        W = sp.csr_matrix((w, c.copy(), D.indptr.copy()), shape=(n, n))
        V = sp.csr_matrix((vv, c.copy(), D.indptr.copy()), shape=(n, n))
        modV = None
        if kappa is not None:
            modV = sp.csr_matrix((v * v * vv / (eps[r] ** 2) / 2.0, c.copy(), D.indptr.copy()), shape=(n, n))
    return Graph(n, knn_ind, knn_dist, D, eps, kappa, W, V, modV)


# --------------------------------------------------------------------------------------
# linear system (GLL.py:29-53) and solvers
# --------------------------------------------------------------------------------------
def laplace_system(W: sp.csr_matrix, Y: np.ndarray, tau: float):
    """L = D - W (degree = column sums, csgraph.laplacian); L_uu + tau I and B = -L_ul Y."""
    n = W.shape[0]
    k_lab = Y.shape[0]
    deg = np.asarray(W.sum(axis=0)).ravel()
    L = (sp.diags(deg) - W).tocsr()
    Luu = (L[k_lab:, k_lab:] + tau * sp.identity(n - k_lab, format="csr")).tocsr()
    B = -(L[k_lab:, :k_lab] @ np.asarray(Y, dtype=np.float64))
    return Luu, B, deg


def textbook_cg(A: sp.csr_matrix, b: np.ndarray, tol: float = 1e-13, max_iter: int = 100000,
                jacobi: bool = True):
    """Multi-RHS (Jacobi-)CG with the per-column freeze of GLL.py:262-269, *without* the
    ``p = r`` aliasing of GLL.py:254.  Returns (x, iterations)."""
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = b[:, None]
    dinv = 1.0 / A.diagonal() if jacobi else np.ones(A.shape[0])
    x = np.zeros_like(b)
    r = b.copy()
    z = r * dinv[:, None]
    p = z.copy()
    rz = np.sum(r * z, axis=0)
    rr = np.sum(r * r, axis=0)
    it = 0
    while np.sqrt(rr.max()) > tol and it < max_iter:
        it += 1
        Ap = A @ p
        live = rr > tol * tol
        pAp = np.sum(p * Ap, axis=0)
        alpha = np.where(live, rz / np.where(live, pAp, 1.0), 0.0)
        x += alpha * p
        r -= alpha * Ap
        z = r * dinv[:, None]
        rz_new = np.sum(r * z, axis=0)
        rr = np.sum(r * r, axis=0)
        live2 = rr > tol * tol
        beta = np.where(live2, rz_new / np.where(rz != 0, rz, 1.0), 0.0)
        p = z + beta * p
        rz = rz_new
    return x, it


def reference_semantics_cg(A, b, x0=None, max_iter=1e5, tol=1e-10):
    """What ``stable_conjgrad`` (GLL.py:247-276) computes, *including* the effect of the
    ``p = r`` alias (the first ``r -= alpha*Ap`` also changes p).  Returns (x, matvecs).
    Used for the CPU "CG iterations/s" baseline and for known-answer tests of the wrapper."""
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
    r = b - A @ x
    p = r  # same object on purpose: this is the behaviour being restated
    rs = np.sum(r * r, axis=0)
    err, it = 1.0, 0
    t2 = tol * tol
    while err > tol and it < max_iter:
        it += 1
        Ap = A @ p
        alpha = np.zeros_like(rs)
        live = rs > t2
        alpha[live] = rs[live] / np.sum(p * Ap, axis=0)[live]
        x += alpha * p
        r -= alpha * Ap
        rs_new = np.sum(r * r, axis=0)
        err = float(np.sqrt(rs_new.max()))
        beta = np.zeros_like(rs)
        live = rs_new > t2
        beta[live] = rs_new[live] / rs[live]
        p = r + beta * p
        rs = rs_new
    return x, it


def solve(Luu: sp.csr_matrix, B: np.ndarray, solver: str = "auto"):
    """GLL.py:53 uses SuperLU; for big systems the oracle uses CG to 1e-13 instead."""
    if solver == "auto":
        solver = "lu" if Luu.shape[0] <= 4000 else "cg"
    if solver == "lu":
        out = spla.spsolve(Luu.tocsc(), B)
        return out.reshape(B.shape)
    x, _ = textbook_cg(Luu, B, tol=1e-13)
    return x.reshape(B.shape)


# --------------------------------------------------------------------------------------
# forward / backward
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class ForwardResult:
    pred: np.ndarray
    graph: Graph
    Luu: sp.csr_matrix
    B: np.ndarray
    deg: np.ndarray


def forward(X, Y, tau: float = 0.0, epsilon: Union[str, float] = "auto", k: int = K_DEFAULT,
            solver: str = "auto", knn: Optional[tuple] = None) -> ForwardResult:
    g = build_graph(X, k, epsilon, knn)
    Luu, B, deg = laplace_system(g.W, Y, tau)
    return ForwardResult(solve(Luu, B, solver), g, Luu, B, deg)


@dataclasses.dataclass
class BackwardResult:
    dX: np.ndarray               # (n,d) float64 (the reference returns fp32)
    w: np.ndarray                # (m,l) adjoint solution
    G: sp.csr_matrix             # per-edge G_ij on the pattern of V
    b: Optional[np.ndarray]      # (n,) adaptive-epsilon row sums, auto only


def backward(X, Y, fwd: ForwardResult, grad_output, solver: str = "auto") -> BackwardResult:
    """GLL.py:93 adjoint solve; GLL.py:104,109 padding; GLL.py:111-120 per-edge
    G_ij = -<w_i - w_j, u_i - u_j>; GLL.py:126-139 adaptive-epsilon term through kappa;
    GLL.py:146-159 out_i = sum_j G_ij V_ij (x_i - x_j) + extra_i."""
    g = fwd.graph
    Xd = np.asarray(X, dtype=np.float64)
    Yd = np.asarray(Y, dtype=np.float64)
    n, k_lab = g.n, Yd.shape[0]
    w = solve(fwd.Luu, np.asarray(grad_output, dtype=np.float64), solver)
    wt = np.concatenate([np.zeros_like(Yd), w], axis=0)
    ut = np.concatenate([Yd, fwd.pred], axis=0)
    indptr, cols = g.V.indptr, g.V.indices
    rows = np.repeat(np.arange(n), np.diff(indptr))
    Gv = -np.einsum("ij,ij->i", wt[rows] - wt[cols], ut[rows] - ut[cols])
    G = sp.csr_matrix((Gv, cols.copy(), indptr.copy()), shape=(n, n))
    coef = Gv * g.V.data
    out = np.zeros_like(Xd)
    # row-wise: out_i = (sum_j c_ij) x_i - sum_j c_ij x_j
    Cm = sp.csr_matrix((coef, cols.copy(), indptr.copy()), shape=(n, n))
    out = np.asarray(Cm.sum(axis=1)).ravel()[:, None] * Xd - Cm @ Xd
    b = None
    if g.kappa is not None:
        b = np.asarray(G.multiply(g.modV).sum(axis=1)).ravel()
        diff = Xd - Xd[g.kappa]
        out -= b[:, None] * diff
        np.add.at(out, g.kappa, b[:, None] * diff)
    return BackwardResult(out, w, G, b)


def fwd_bwd(X, Y, y_query, tau=0.0, epsilon="auto", k=K_DEFAULT, solver="auto"):
    """One benchmark 'call' (SURVEY.md 8d): pred = layer(X,Y); loss = custom_ce_loss; backward."""
    f = forward(X, Y, tau, epsilon, k, solver)
    loss, gout = ce_loss_and_grad(f.pred, y_query)
    bw = backward(X, Y, f, gout, solver)
    return f, loss, gout, bw


# --------------------------------------------------------------------------------------
# comparison helpers shared by the parity tests
# --------------------------------------------------------------------------------------
def max_rel(a, ref) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    return float(np.max(np.abs(a - ref)) / max(np.max(np.abs(ref)), 1e-300))


def knn_sets_match(ind, ref_ind, ref_dist, rtol: float = 1e-6):
    """Per row set equality; a swap is tolerated only between entries whose oracle distance
    is within rtol (relative) of the row's k-th distance (documented near-tie).
    Returns (n_exact_rows, n_tie_rows, n_bad_rows)."""
    n = ref_ind.shape[0]
    a = np.sort(np.asarray(ind, dtype=np.int64), axis=1)
    b = np.sort(np.asarray(ref_ind, dtype=np.int64), axis=1)
    same = np.all(a == b, axis=1)
    tie = bad = 0
    for i in np.nonzero(~same)[0]:
        missing = np.setdiff1d(ref_ind[i], ind[i])
        dk = ref_dist[i, -1]
        pos = {int(j): t for t, j in enumerate(ref_ind[i])}
        if all(abs(ref_dist[i, pos[int(j)]] - dk) <= rtol * max(dk, 1e-30) for j in missing):
            tie += 1
        else:
            bad += 1
    return int(same.sum()), tie, bad
