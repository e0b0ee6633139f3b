"""CPU tests: pin the restated oracle (oracle/gll_oracle.py) against the fixtures produced by the
UNMODIFIED reference GLL.py (oracle/make_golden.py), and check the invariants of SURVEY.md 4."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_names, load_golden
from oracle import gll_oracle as O

SMALL = [n for n in golden_names() if n.startswith("g_") or n.startswith("c1_")]
BIG = [n for n in golden_names() if n.startswith(("c2_", "c3_"))]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_fixture(name):
    g, X, Y, yq = load_golden(name)
    f, loss, gout, bw = O.fwd_bwd(X, Y, yq, g["tau_arg"], g["eps_arg"], solver="lu")
    s = int(g["dX_stride"])
    # kNN lists: identical to what the shim gave the reference (independent implementations)
    assert np.array_equal(f.graph.knn_ind[::s], g["knn_ind"])
    assert np.array_equal(f.graph.knn_dist[::s].astype(np.float32), g["knn_dist"])
    assert O.max_rel(f.pred, g["pred"]) < 1e-9
    assert abs(loss - float(g["loss"])) < 1e-9
    # the reference casts the Laplacian to fp32 and multiplies in fp32 (GLL.py:134,154)
    assert O.max_rel(bw.dX[::s], g["dX"]) < 5e-6
    assert abs(np.linalg.norm(bw.dX) / float(g["dX_fro"]) - 1) < 1e-5


@pytest.mark.parametrize("name", BIG)
def test_oracle_matches_reference_fixture_big(name):
    g, X, Y, yq = load_golden(name)
    f, loss, gout, bw = O.fwd_bwd(X, Y, yq, g["tau_arg"], g["eps_arg"], solver="lu")
    s = int(g["dX_stride"])
    assert np.array_equal(f.graph.knn_ind[::s], g["knn_ind"])
    assert O.max_rel(f.pred, g["pred"]) < 1e-9
    assert O.max_rel(bw.dX[::s], g["dX"]) < 5e-6


def test_invariants_tau0_rows_sum_to_one_and_range():
    X, Y, _, yq = O.synth_inputs(7, 200, 300, 32, 5, 1.5)
    f = O.forward(X, Y, 0.0, "auto")
    assert np.allclose(f.pred.sum(axis=1), 1.0, atol=1e-12)
    assert f.pred.min() >= -1e-12 and f.pred.max() <= 1 + 1e-12
    f2 = O.forward(X, Y, 0.07, 1.0)
    assert np.all(f2.pred.sum(axis=1) < 1.0)
    W = f.graph.W
    assert abs(W - W.T).max() < 1e-15
    assert W.diagonal().max() == 0.0
    assert np.diff(W.indptr).min() >= 24


@pytest.mark.parametrize("eps,tau", [(1.0, 0.07), ("auto", 0.0), ("auto", 0.07)])
def test_backward_matches_finite_differences(eps, tau):
    X, Y, _, yq = O.synth_inputs(11, 120, 180, 24, 4, 1.5)
    f, loss, gout, bw = O.fwd_bwd(X, Y, yq, tau, eps, solver="lu")
    knn = (f.graph.knn_ind, None)
    rng = np.random.default_rng(0)
    Xd = X.astype(np.float64)

    def loss_at(Xp):
        # hold the kNN *sets* fixed (the layer is only piecewise differentiable), recompute
        # distances / eps / weights from the perturbed features
        ind = f.graph.knn_ind
        dist = np.sqrt(((Xp[:, None, :] - Xp[ind]) ** 2).sum(-1))
        dist[:, 0] = 0.0
        ff = O.forward(Xp, Y, tau, eps, knn=(ind, dist), solver="lu")
        return O.ce_loss_and_grad(ff.pred, yq)[0]

    for _ in range(3):
        dirn = rng.standard_normal(Xd.shape)
        h = 1e-6
        fd = (loss_at(Xd + h * dirn) - loss_at(Xd - h * dirn)) / (2 * h)
        an = float(np.sum(bw.dX * dirn))
        assert abs(fd - an) <= 2e-4 * max(abs(an), 1e-6), (fd, an)
    assert np.linalg.norm(bw.dX[:120]) > 0.05 * np.linalg.norm(bw.dX)  # base rows get gradient too


def test_cg_variants_agree_with_direct_solve():
    X, Y, _, _ = O.synth_inputs(3, 150, 450, 16, 6, 1.0)
    f = O.forward(X, Y, 0.05, 1.0, solver="lu")
    x_cg, it = O.textbook_cg(f.Luu, f.B, tol=1e-12)
    assert O.max_rel(x_cg, f.pred) < 1e-9
    x_ref, it_ref = O.reference_semantics_cg(f.Luu, f.B, tol=1e-10)
    assert O.max_rel(x_ref, f.pred) < 1e-7
    res = np.sqrt(((f.Luu @ x_ref - f.B) ** 2).sum(axis=0)).max()
    assert res < 1e-9
    assert it_ref >= it  # the p = r alias of GLL.py:254 costs iterations, never accuracy


def test_knn_tie_helper():
    ref_ind = np.array([[0, 1, 2, 3]])
    ref_dist = np.array([[0.0, 0.5, 1.0, 1.0 + 5e-7]])
    assert O.knn_sets_match(np.array([[0, 1, 2, 3]]), ref_ind, ref_dist) == (1, 0, 0)
    assert O.knn_sets_match(np.array([[0, 1, 2, 9]]), ref_ind, ref_dist) == (0, 1, 0)
    assert O.knn_sets_match(np.array([[0, 9, 2, 3]]), ref_ind, ref_dist) == (0, 0, 1)


def _load_eval_fixture():
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_k50.npz"))
    seed, k_lab, m, d, l, knn = (int(v) for v in g["params"])
    X, Y, _, yq = O.synth_inputs(seed, k_lab, m, d, l, float(g["sigma"]))
    import hashlib

    assert hashlib.sha256(X.tobytes()).hexdigest() == str(g["x_sha256"])
    W = sp.csr_matrix((g["w_data"], g["w_indices"], g["w_indptr"]), shape=(k_lab + m, k_lab + m))
    return g, X, Y, yq, k_lab, knn, W


def test_oracle_matches_reference_eval_path_fixture():
    """The reference's evaluation routine (utils.py:570-593) on the UNMODIFIED knn_sym_dist (k = 50) and stable_conjgrad of
    GLL.py (oracle/make_golden_eval.py): the restated graph and a textbook CG on the same Jacobi-scaled system agree."""
    g, X, Y, yq, k_lab, knn, W_ref = _load_eval_fixture()
    gr = O.build_graph(X, knn, "auto")
    W = sp.csr_matrix(gr.W)
    W.sort_indices()
    assert np.array_equal(W.indptr, W_ref.indptr) and np.array_equal(W.indices, W_ref.indices)
    assert np.max(np.abs(W.data - W_ref.data)) < 1e-12
    L = (sp.diags(np.asarray(W.sum(axis=0)).ravel()) - W).tocsr()
    Luu = (L[k_lab:, k_lab:] + float(g["tau"]) * sp.identity(X.shape[0] - k_lab)).tocsr()
    M = sp.diags(1.0 / np.sqrt(Luu.diagonal() + 1e-10))
    y = O.textbook_cg(sp.csr_matrix(M @ Luu @ M), -(M @ (L[k_lab:, :k_lab] @ Y.astype(np.float64))), tol=1e-12)[0]
    assert O.max_rel(M @ y, g["pred"]) < 1e-8   # the reference stops at a 1e-10 residual
