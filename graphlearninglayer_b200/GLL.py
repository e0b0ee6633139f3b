"""Host side of the B200 GraphLearningLayer hot path: the same three names the reference module exports
(``/root/reference/GLL.py``; ``utils.py:25`` imports all three), backed by libgll_b200.so.

    LaplaceLearningSparseHard.apply(X, label_matrix, tau=0, epsilon='auto')   GLL.py:10-177
    knn_sym_dist(data, k=25, epsilon='auto')                                  GLL.py:180-244
    stable_conjgrad(A, b, x0=None, max_iter=1e5, tol=1e-10)                   GLL.py:247-276

PyTorch is plumbing here (device memory, the current CUDA stream, autograd registration); every number is
produced by the sm_100a kernels in csrc/.  There is no CPU path: CPU tensors raise.

Knobs (environment, read at call time):
    GLL_B200_CG_TOL       absolute 2-norm residual per class column (default 1e-7; forward), and relative to the
                          largest column norm of grad_output for the adjoint solve
    GLL_B200_CG_MAXIT     default 5000
    GLL_B200_PRED_DTYPE   'float64' (default: the reference returns float64, GLL.py:66) or 'float32'
    GLL_B200_CHECK        '1': read the device status word after every call (one host sync) and warn like the
                          reference does (GLL.py:240-241 epsilon ~ 0; GLL.py:273-274 'max iter reached').
                          Default: DEFERRED -- the status word is copied to pinned memory behind the call and looked at
                          when a later call finds the copy complete (no host sync; the warning comes one call late).
                          '0': never (also skipped while a CUDA graph is being captured).
"""
from __future__ import annotations

import os
import warnings
from typing import Optional, Union

import numpy as np
import torch

from . import _lib
from ._lib import lib

K_NEIGHBOURS = 25  # hard-coded in the reference at GLL.py:27

__all__ = ["LaplaceLearningSparseHard", "LaplaceLearningSparseHardNormalized", "knn_sym_dist", "stable_conjgrad", "last_info"]


def _env_float(name: str, default: float) -> float:
    v = os.environ.get(name)
    return float(v) if v else default


def _cg_tol() -> float:
    return _env_float("GLL_B200_CG_TOL", 1e-7)


def _cg_maxit() -> int:
    return int(_env_float("GLL_B200_CG_MAXIT", 5000))


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _bytes(nbytes: int, device: torch.device) -> torch.Tensor:
    # PyTorch's caching allocator owns every byte the library touches (the library never allocates).
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"graphlearninglayer_b200: {what} must be a CUDA tensor (this build has no CPU path; "
                           "the reference /root/reference/GLL.py is the CPU implementation)")


_last_info: Optional[torch.Tensor] = None


def last_info() -> dict:
    """Device status block of the most recent forward/backward (synchronises).  Keys follow GLL_INFO_* in
    include/gll_b200.h."""
    if _last_info is None:
        return {}
    v = _last_info.cpu().numpy()
    f = v.view(np.float32)
    return dict(status=int(v[_lib.INFO_STATUS]), nnz=int(v[_lib.INFO_NNZ]), nnz_uu=int(v[_lib.INFO_NNZ_UU]),
                cg_iters_fwd=int(v[_lib.INFO_CG_ITERS_FWD]), cg_iters_bwd=int(v[_lib.INFO_CG_ITERS_BWD]),
                knn_fallback_rows=int(v[_lib.INFO_KNN_FALLBACK_ROWS]), cg_resid_fwd=float(f[_lib.INFO_CG_RESID_FWD]),
                cg_resid_bwd=float(f[_lib.INFO_CG_RESID_BWD]))


def _warn_from_status(info: torch.Tensor, where: str) -> None:
    st = int(info[_lib.INFO_STATUS].item())
    if st & _lib.STATUS_EPS_TINY:
        warnings.warn("Epsilon in KNN is very close to zero.")  # GLL.py:240-241
    if st & _lib.STATUS_CG_NOT_CONVERGED:
        warnings.warn(f"max iter reached in the {where} CG solve")  # GLL.py:273-274 prints
    if st & _lib.STATUS_NONFINITE:
        warnings.warn(f"non-finite values in the {where} solve (singular L_uu or epsilon = 0)")


class _DeferredStatus:
    """One pinned slot per device: the previous call's status word, copied asynchronously; read when its event has completed."""
    slots: dict = {}

    def __init__(self):
        self.host = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32).pin_memory()
        self.event: Optional[torch.cuda.Event] = None
        self.where = ""


def _check_status(info: torch.Tensor, where: str) -> None:
    mode = os.environ.get("GLL_B200_CHECK", "")
    if mode == "1":
        _warn_from_status(info, where)
        return
    if mode == "0" or torch.cuda.is_current_stream_capturing():
        return
    key = (info.device.index, torch.cuda.current_stream(info.device).cuda_stream)
    slot = _DeferredStatus.slots.get(key)
    if slot is None:
        slot = _DeferredStatus.slots[key] = _DeferredStatus()
    if slot.event is not None:
        if not slot.event.query():
            return  # the previous copy has not landed yet: keep it, look again at the next call
        _warn_from_status(slot.host, slot.where)
    slot.host.copy_(info, non_blocking=True)
    slot.where = where
    slot.event = torch.cuda.Event()
    slot.event.record(torch.cuda.current_stream(info.device))


class _State:
    """Buffers kept between forward and backward (replaces the scipy objects on ctx, GLL.py:69-70)."""
    __slots__ = ("buf", "layout", "n", "d", "k", "l", "k_lab", "eps_auto")

    def view(self, name: str, dtype: torch.dtype, count: int) -> torch.Tensor:
        off = getattr(self.layout, name)
        item = torch.empty((), dtype=dtype).element_size()
        return self.buf[off:off + count * item].view(dtype)


def _forward_impl(X: torch.Tensor, label_matrix: torch.Tensor, tau, epsilon, k: int = K_NEIGHBOURS):
    _require_cuda(X, "features")
    if X.dim() != 2:
        raise ValueError("features must be (n, d)")
    dev = X.device
    Xc = X.detach()
    if Xc.dtype != torch.float32:
        Xc = Xc.float()
    Xc = Xc.contiguous()
    Y = label_matrix.detach().to(device=dev, dtype=torch.float32).contiguous()  # int64 one-hot allowed (t_a_a.py:545)
    if Y.dim() != 2:
        raise ValueError("label_matrix must be (k_lab, l)")
    n, d = Xc.shape
    k_lab, l = Y.shape
    if not (0 < k_lab < n):
        raise ValueError(f"need 0 < k_lab < n (k_lab={k_lab}, n={n}); labeled rows come first (GLL.py:11)")
    if n < k:
        raise ValueError(f"need at least k={k} nodes, got {n}")
    if isinstance(epsilon, str):
        if epsilon != "auto":
            raise ValueError("epsilon must be a float or 'auto'")
        eps_auto, eps_fixed = 1, 0.0
    else:
        eps_auto, eps_fixed = 0, float(epsilon)
    m = n - k_lab
    st = _State()
    st.layout = _lib.state_layout(n, k, l, k_lab)
    st.n, st.d, st.k, st.l, st.k_lab, st.eps_auto = n, d, k, l, k_lab, eps_auto
    pred64 = os.environ.get("GLL_B200_PRED_DTYPE", "float64") != "float32"
    with torch.cuda.device(dev):
        st.buf = _bytes(st.layout.total, dev)
        ws_bytes = lib.gll_workspace_bytes(n, d, k, l, k_lab)
        ws = _bytes(ws_bytes, dev)
        pred = torch.empty((m, l), dtype=torch.float64 if pred64 else torch.float32, device=dev)
        rc = lib.gll_forward(Xc.data_ptr(), Y.data_ptr(), n, d, k, l, k_lab, eps_auto, eps_fixed, float(tau), _cg_tol(),
                             _cg_maxit(), st.buf.data_ptr(), pred.data_ptr(), int(pred64), ws.data_ptr(), ws_bytes,
                             _stream_ptr(dev))
    _lib.check(rc, "gll_forward")
    global _last_info
    _last_info = st.view("info", torch.int32, _lib.INFO_WORDS)
    _check_status(_last_info, "forward")
    return pred, st, Xc


def _backward_impl(st: _State, Xc: torch.Tensor, grad_output: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dX for d loss / d Pred = grad_output (times the 0-dim device tensor `scale`, if given: read by the kernel, no host sync)."""
    dev = Xc.device
    g = grad_output.detach().to(device=dev)
    if g.dtype not in (torch.float32, torch.float64):
        g = g.float()
    g = g.contiguous()
    m = st.n - st.k_lab
    if tuple(g.shape) != (m, st.l):
        raise ValueError(f"grad_output must be ({m}, {st.l}), got {tuple(g.shape)}")
    with torch.cuda.device(dev):
        dX = torch.empty((st.n, st.d), dtype=torch.float32, device=dev)
        ws_bytes = lib.gll_workspace_bytes(st.n, st.d, st.k, st.l, st.k_lab)
        ws = _bytes(ws_bytes, dev)
        # negative tolerance = relative to the largest column norm of the right-hand side (see gll_cg_solve)
        sc = None
        if scale is not None:
            sc = scale.detach().to(device=dev)
            if sc.dtype not in (torch.float32, torch.float64):
                sc = sc.float()
            sc = sc.reshape(1).contiguous()
        rc = lib.gll_backward_scaled(Xc.data_ptr(), g.data_ptr(), int(g.dtype == torch.float64),
                                     sc.data_ptr() if sc is not None else None, int(sc is not None and sc.dtype == torch.float64),
                                     st.n, st.d, st.k, st.l, st.k_lab, st.eps_auto, -_cg_tol(), _cg_maxit(), st.buf.data_ptr(),
                                     dX.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(dev))
    _lib.check(rc, "gll_backward")
    _check_status(st.view("info", torch.int32, _lib.INFO_WORDS), "adjoint")
    return dX


class LaplaceLearningSparseHard(torch.autograd.Function):
    """Drop-in for ``GLL.LaplaceLearningSparseHard`` (GLL.py:10-177).

    forward(X, label_matrix, tau=0, epsilon='auto'): X is (n, d) with the k_lab labeled ("base") rows FIRST
    (GLL.py:11,32-38), label_matrix is (k_lab, l) one-hot (float32 or int64); returns the (n-k_lab, l) harmonic
    extension, float64 like the reference (GLL.py:66,73), as a fresh tensor callers may mutate (adversarial.py:691).
    backward returns (dX, None, None, None) with dX (n, d) in X's dtype, non-zero on base rows too (GLL.py:159,177).
    """

    @staticmethod
    def forward(ctx, X, label_matrix, tau=0, epsilon="auto"):
        pred, st, Xc = _forward_impl(X, label_matrix, tau, epsilon)
        ctx.gll_state = st
        ctx.x_dtype = X.dtype
        ctx.save_for_backward(Xc)
        ctx.set_materialize_grads(True)
        return pred

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        (Xc,) = ctx.saved_tensors
        dX = _backward_impl(ctx.gll_state, Xc, grad_output)
        if dX.dtype != ctx.x_dtype:
            dX = dX.to(ctx.x_dtype)
        return dX, None, None, None


class LaplaceLearningSparseHardNormalized(torch.autograd.Function):
    """`LaplaceLearningSparseHard.apply(F.normalize(feat, dim=1), label_matrix, tau, epsilon)` in one autograd node
    (SURVEY 8f-3): every caller normalises the encoder output right before the layer (networks/BuildNet.py:101,
    FullySup.py:122,156).  The rows are normalised by one kernel and the normalisation's backward,
    d feat = (d xn - xn <xn, d xn>) / |feat|, by another -- instead of PyTorch's ~10 small kernels for the pair."""

    @staticmethod
    def forward(ctx, feat, label_matrix, tau=0, epsilon="auto"):
        _require_cuda(feat, "features")
        f = feat.detach()
        if f.dtype != torch.float32:
            f = f.float()
        f = f.contiguous()
        n, d = f.shape
        xn = torch.empty_like(f)
        inv = torch.empty(n, dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            _lib.check(lib.gll_normalize_rows(f.data_ptr(), n, d, 1e-12, xn.data_ptr(), inv.data_ptr(), _stream_ptr(f.device)),
                       "gll_normalize_rows")
        pred, st, _ = _forward_impl(xn, label_matrix, tau, epsilon)
        ctx.gll_state = st
        ctx.x_dtype = feat.dtype
        ctx.save_for_backward(xn, inv)
        ctx.set_materialize_grads(True)
        return pred

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        xn, inv = ctx.saved_tensors
        dxn = _backward_impl(ctx.gll_state, xn, grad_output)
        dx = torch.empty_like(dxn)
        n, d = xn.shape
        with torch.cuda.device(xn.device):
            _lib.check(lib.gll_normalize_rows_backward(xn.data_ptr(), inv.data_ptr(), dxn.data_ptr(), n, d, dx.data_ptr(),
                                                       _stream_ptr(xn.device)), "gll_normalize_rows_backward")
        if dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        return dx, None, None, None


# ------------------------------------------------------------------------------------------------------------
# numpy / scipy compatibility wrappers used by the reference's evaluation code (utils.py:570-593)
# ------------------------------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("graphlearninglayer_b200 needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _graph_stages(Xc: torch.Tensor, k: int, epsilon):
    """kNN -> CSR -> weights through the stage entry points; returns device tensors."""
    dev = Xc.device
    n, d = Xc.shape
    emax = lib.gll_max_edges(n, k)
    i32, f32 = torch.int32, torch.float32
    knn_idx = torch.empty((n, k), dtype=i32, device=dev)
    knn_dist = torch.empty((n, k), dtype=f32, device=dev)
    row_ptr = torch.empty(n + 1, dtype=i32, device=dev)
    col = torch.empty(emax, dtype=i32, device=dev)
    dist = torch.empty(emax, dtype=f32, device=dev)
    info = torch.zeros(_lib.INFO_WORDS, dtype=i32, device=dev)
    s = _stream_ptr(dev)
    wsb = max(lib.gll_knn_workspace_bytes(n, d, k), lib.gll_graph_workspace_bytes(n, k), lib.gll_weights_workspace_bytes(n, k))
    ws = _bytes(wsb, dev)
    _lib.check(lib.gll_knn(Xc.data_ptr(), n, d, k, knn_idx.data_ptr(), knn_dist.data_ptr(), info.data_ptr(), ws.data_ptr(),
                           wsb, s), "gll_knn")
    _lib.check(lib.gll_graph_build(knn_idx.data_ptr(), knn_dist.data_ptr(), n, k, row_ptr.data_ptr(), col.data_ptr(),
                                   dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "gll_graph_build")
    eps_auto = isinstance(epsilon, str)
    if eps_auto and epsilon != "auto":
        raise ValueError("epsilon must be a float or 'auto'")
    # the weights stage also emits the unlabeled-block system; with a 1-row dummy label block it is just ignored
    k_lab, l = 1, 1
    lp = lib.gll_padded_classes(l)
    m = n - k_lab
    Y = torch.zeros((k_lab, l), dtype=f32, device=dev)
    out = dict(eps=torch.empty(n, dtype=f32, device=dev), kappa=torch.empty(n, dtype=i32, device=dev),
               w=torch.empty(emax, dtype=f32, device=dev), deg=torch.empty(n, dtype=f32, device=dev))
    scratch = dict(uu_ptr=torch.empty(m + 1, dtype=i32, device=dev), uu_col=torch.empty(emax, dtype=i32, device=dev),
                   uu_val=torch.empty(emax, dtype=f32, device=dev), diag=torch.empty(m, dtype=f32, device=dev),
                   rhs=torch.empty(m * lp, dtype=f32, device=dev), ut=torch.empty(n * lp, dtype=f32, device=dev))
    _lib.check(lib.gll_edge_weights(knn_idx.data_ptr(), knn_dist.data_ptr(), row_ptr.data_ptr(), col.data_ptr(),
                                    dist.data_ptr(), Y.data_ptr(), n, k, l, k_lab, int(eps_auto),
                                    0.0 if eps_auto else float(epsilon), 0.0, out["eps"].data_ptr(),
                                    out["kappa"].data_ptr(), out["w"].data_ptr(), out["deg"].data_ptr(),
                                    scratch["uu_ptr"].data_ptr(), scratch["uu_col"].data_ptr(),
                                    scratch["uu_val"].data_ptr(), scratch["diag"].data_ptr(), scratch["rhs"].data_ptr(),
                                    scratch["ut"].data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "gll_edge_weights")
    return dict(knn_idx=knn_idx, knn_dist=knn_dist, row_ptr=row_ptr, col=col, dist=dist, info=info, **out)


def knn_sym_dist(data, k=K_NEIGHBOURS, epsilon="auto"):
    """Same contract as the reference ``knn_sym_dist`` (GLL.py:180-244): numpy (n, d) in, scipy CSR out.

    Returns ``(W, V, mod_V, C, knn_ind)``; for a fixed epsilon ``mod_V`` and ``C`` are the integer 0, the reference's
    placeholders (GLL.py:229-230; its backward tests ``isinstance(C, int)``, GLL.py:124).  For 'auto', ``C`` is
    returned as a sparse CSR with C[kappa(i), i] = 1 (the reference builds the same matrix through a dense n x n
    array, GLL.py:209-213).  kNN, symmetrisation and the exp() run on the GPU; V and mod_V are the closed-form
    multiples of W (GLL.py:217-218) formed in fp64 on the host for the caller.
    """
    import scipy.sparse as sp

    dev = _device()
    Xc = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)).to(dev)
    g = _graph_stages(Xc, int(k), epsilon)
    n = Xc.shape[0]
    indptr = g["row_ptr"].cpu().numpy()
    E = int(indptr[-1])
    cols = g["col"][:E].cpu().numpy()
    dist = g["dist"][:E].cpu().numpy().astype(np.float64)
    w = g["w"][:E].cpu().numpy().astype(np.float64)
    eps = g["eps"].cpu().numpy().astype(np.float64)
    if int(g["info"][_lib.INFO_STATUS].item()) & _lib.STATUS_EPS_TINY:
        warnings.warn("Epsilon in KNN is very close to zero.")  # GLL.py:240-241
    rows = np.repeat(np.arange(n), np.diff(indptr))
    with np.errstate(divide="ignore", invalid="ignore"):
        v = -8.0 * w / eps[rows] / eps[cols]
    W = sp.csr_matrix((w, cols, indptr), shape=(n, n))
    V = sp.csr_matrix((v, cols.copy(), indptr.copy()), shape=(n, n))
    knn_ind = g["knn_idx"].cpu().numpy().astype(np.int64)
    if isinstance(epsilon, str):
        with np.errstate(divide="ignore", invalid="ignore"):
            mv = dist * dist * v / (eps[rows] ** 2) / 2.0
        mod_V = sp.csr_matrix((mv, cols.copy(), indptr.copy()), shape=(n, n))
        kappa = g["kappa"].cpu().numpy().astype(np.int64)
        Cm = sp.csr_matrix((np.ones(n), (kappa, np.arange(n))), shape=(n, n))
        return W, V, mod_V, Cm, knn_ind
    return W, V, 0, 0, knn_ind  # placeholders of GLL.py:229-230


def stable_conjgrad(A, b, x0=None, max_iter=1e5, tol=1e-10):
    """Same contract as the reference ``stable_conjgrad`` (GLL.py:247-276): scipy sparse SPD ``A``, numpy ``b``
    (m,) or (m, l) -> numpy solution with max_c ||b_c - A x_c||_2 <= tol (absolute), or 'max iter reached'.

    The solves run in the persistent fp32 CG kernel (gll_cg_solve); tolerances below what fp32 can reach are met by
    fp64 iterative refinement around it (residual b - A x by an fp64 CSR kernel, correction solved by the CG kernel), so
    the reference's default tol=1e-10 (utils.py:589-591) is honoured.  The ``p = r`` alias of GLL.py:254 is not
    reproduced (it only costs the reference iterations).
    """
    import scipy.sparse as sp

    dev = _device()
    A = sp.csr_matrix(A)
    A.sort_indices()
    b_np = np.asarray(b, dtype=np.float64)
    one_d = b_np.ndim == 1
    B = b_np[:, None] if one_d else b_np
    m, l_all = B.shape
    dg = A.diagonal()
    off = (A - sp.diags(dg)).tocsr()
    off.eliminate_zeros()
    off.sort_indices()
    i32, f32, f64 = torch.int32, torch.float32, torch.float64
    ptr = torch.as_tensor(off.indptr.astype(np.int32)).to(dev)
    col = torch.as_tensor(off.indices.astype(np.int32)).to(dev)
    val = torch.as_tensor((-off.data).astype(np.float32)).to(dev)  # kernel applies diag*p - sum val*p
    diag = torch.as_tensor(dg.astype(np.float32)).to(dev)
    # fp64 copy of the full matrix for the refinement residual r = b - A x (gll_csr_residual_f64)
    a_ptr = torch.as_tensor(A.indptr.astype(np.int32)).to(dev)
    a_col = torch.as_tensor(A.indices.astype(np.int32)).to(dev)
    a_val = torch.as_tensor(A.data.astype(np.float64)).to(dev)

    def residual(bb, x):
        r = torch.empty_like(bb)
        _lib.check(lib.gll_csr_residual_f64(a_ptr.data_ptr(), a_col.data_ptr(), a_val.data_ptr(), x.data_ptr(), bb.data_ptr(), m,
                                            bb.shape[1], r.data_ptr(), _stream_ptr(dev)), "gll_csr_residual_f64")
        return r

    X = np.zeros_like(B) if x0 is None else np.array(np.asarray(x0, dtype=np.float64).reshape(B.shape))
    maxit = int(max_iter)
    s = _stream_ptr(dev)
    total_it = 0
    converged = True
    for c0 in range(0, l_all, 128):  # the kernel handles up to 128 class columns per launch
        c1 = min(l_all, c0 + 128)
        l = c1 - c0
        lp = lib.gll_padded_classes(l)
        wsb = lib.gll_cg_workspace_bytes(m, l)
        ws = _bytes(wsb, dev)
        x = torch.as_tensor(X[:, c0:c1]).to(dev, f64).contiguous()
        bb = torch.as_tensor(B[:, c0:c1]).to(dev, f64).contiguous()
        rhs = torch.zeros((m, lp), dtype=f32, device=dev)
        corr = torch.empty((m, lp), dtype=f32, device=dev)
        stat = torch.zeros(4, dtype=i32, device=dev)
        ok = False
        for _ in range(12):
            r = residual(bb, x)
            rn = float(torch.linalg.vector_norm(r, dim=0).max().item())
            if rn <= tol or total_it >= maxit:
                ok = rn <= tol
                break
            scale = rn
            rhs[:, :l] = (r / scale).to(f32)
            inner_tol = max(1e-6, min(1e-3, 0.1 * tol / scale))
            _lib.check(lib.gll_cg_solve(ptr.data_ptr(), col.data_ptr(), val.data_ptr(), diag.data_ptr(), rhs.data_ptr(), m, l,
                                        inner_tol, max(1, maxit - total_it), corr.data_ptr(), stat.data_ptr(), 0,
                                        stat[1:].data_ptr(), ws.data_ptr(), wsb, s), "gll_cg_solve")
            total_it += int(stat[0].item())
            x += corr[:, :l].to(f64) * scale
        else:
            r = residual(bb, x)
            ok = float(torch.linalg.vector_norm(r, dim=0).max().item()) <= tol
        converged &= ok
        X[:, c0:c1] = x.cpu().numpy()
    if not converged:
        print("max iter reached")  # GLL.py:273-274
    return X[:, 0] if one_d else X
