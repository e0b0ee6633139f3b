"""CPU tests of the error bound behind the tensor-core kNN search's completeness proof (oracle/split_model.py restates the
fp16 operand split of csrc/knn.cu and the bound of knn_tc_err_coef / knn_err_bound): the approximate squared distances
may only be off by the bound, for any feature scale -- otherwise the exact re-rank could miss a neighbour silently."""
import numpy as np
import pytest

from oracle import gll_oracle as O
from oracle import split_model as S


def _datasets():
    rng = np.random.default_rng(0)
    X, *_ = O.synth_inputs(3, 300, 340, 512, 10, 4.5)           # the C2 distribution (unit rows)
    yield "unit_rows_d512", X
    X2, *_ = O.synth_inputs(4, 200, 300, 200, 10, 3.0)          # d not a multiple of 16
    yield "unit_rows_d200", X2
    yield "scaled_3e4", (X2 * np.float32(3.0e4)).astype(np.float32)
    yield "scaled_2e-6", (X2 * np.float32(2.0e-6)).astype(np.float32)
    yield "ragged_norms", (X2 * np.logspace(-3, 3, X2.shape[0], dtype=np.float32)[rng.permutation(X2.shape[0]), None]).astype(np.float32)
    relu = np.maximum(rng.standard_normal((400, 256)), 0).astype(np.float32) ** 3   # sparse, heavy-tailed (post-ReLU-like)
    yield "relu_cubed", relu
    spike = (1e-6 * rng.standard_normal((300, 128))).astype(np.float32)             # one dominant element: the rest goes subnormal in fp16
    spike[np.arange(300), rng.integers(0, 128, 300)] = 1.0
    yield "spike_plus_dust", spike
    dup = rng.standard_normal((200, 64)).astype(np.float32)
    dup[50:120] = dup[50]                                                            # duplicates: exact distance 0
    dup[7] = 0.0                                                                     # a zero row
    yield "duplicates_and_zero_row", dup


@pytest.mark.parametrize("name,X", list(_datasets()), ids=[n for n, _ in _datasets()])
@pytest.mark.parametrize("fp32_acc,flush", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("passes", [1, 2])
def test_f16_distance_error_within_proven_bound(name, X, fp32_acc, flush, passes):
    """One pass (hi.hi, the default) and two passes ((hi + lo).hi).  flush: the same with every fp16 subnormal operand
    replaced by zero -- the bound does not lean on how the tensor core treats subnormals (rows are scaled to a norm of ~2^8,
    so only elements below 2^-22 of the row norm are subnormal in hi)."""
    hi, lo, E, sq, rho = S.split_f16(X)
    d = X.shape[1]
    approx = S.approx_d2_f16(hi, lo, E, sq, passes=passes, fp32_accumulate=fp32_acc, flush_subnormals=flush).astype(np.float64)
    exact = S.exact_d2(X)
    err = np.abs(approx - exact)
    bound = S.err_bound(d, sq, rho, passes)[:, None]
    worst = float((err / np.maximum(bound, 1e-300)).max())
    assert worst <= 1.0, (name, worst)
    # ... and the bound is not vacuous: for unit rows it stays below the 24th -> 32nd neighbour gap of the benchmark graphs (3e-3)
    if name.startswith("unit_rows"):
        assert float(bound.max()) < (1.5e-3 if passes == 1 else 1.0e-3)
        assert float(err.max()) < 0.5 * float(bound.max())   # what the Gram actually does: well inside its rigorous bound


def test_rows_normalised_to_one_share_one_scale():
    """Every caller normalises the features (networks/BuildNet.py:101): all rows must land in the E = 0 bucket whatever their
    rounding, because the Gram epilogue then takes its single-FMA path."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((5000, 96)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True).astype(np.float32)
    E = S.scale_exponent((X.astype(np.float64) ** 2).sum(axis=1))
    assert (E == -S.F16_TARGET_LOG2).all()
    # bucket edges: 2^(2E-1) <= 1.5 |x|^2 < 2^(2E+1)
    s = np.array([0.0, 1e-30, 1 / 3 - 1e-6, 1 / 3 + 1e-6, 4 / 3 - 1e-6, 4 / 3 + 1e-6, 16 / 3 + 1e-5, 1e30, np.inf])
    assert (S.scale_exponent(s) + S.F16_TARGET_LOG2).tolist() == [0, -50, -1, 0, 0, 1, 2, 50, 0]


def test_scaled_rows_fit_fp16_without_overflow():
    rng = np.random.default_rng(2)
    X = (rng.standard_normal((100, 40)) * 10.0 ** rng.uniform(-15, 15, (100, 1))).astype(np.float32)
    hi, lo, E, sq, rho = S.split_f16(X)
    assert np.isfinite(hi.astype(np.float32)).all() and np.isfinite(lo.astype(np.float32)).all()
    z = np.abs(hi.astype(np.float64))
    assert z.max() < 1.16 * 256 and np.linalg.norm(hi.astype(np.float64), axis=1).min() > 0.57 * 256


def _hard_datasets():
    rng = np.random.default_rng(5)
    X, *_ = O.synth_inputs(9, 300, 500, 128, 10, 3.0)
    yield "clusters_d128", X
    tight = (1.0 + 2e-3 * rng.standard_normal((600, 48))).astype(np.float32)   # one tight blob: neighbour gaps far below the bound
    yield "tight_blob", tight
    grid = np.stack(np.meshgrid(np.arange(25.0), np.arange(24.0)), -1).reshape(-1, 2).astype(np.float32)  # lattice: many exact ties
    yield "lattice_ties", np.hstack([grid, np.zeros((grid.shape[0], 6), np.float32)])
    dup = rng.standard_normal((300, 32)).astype(np.float32)
    dup[100:160] = dup[100]
    yield "duplicates", dup


@pytest.mark.parametrize("name,X", list(_hard_datasets()), ids=[n for n, _ in _hard_datasets()])
@pytest.mark.parametrize("passes", [1, 2])
def test_proven_rows_equal_exact_knn(name, X, passes):
    """Selection by approximate distance + exact re-rank + completeness proof (the GPU pipeline, restated): every row the
    proof accepts must carry exactly the oracle's neighbour list; rows it rejects are the fallback's business.  On well
    separated data almost every row is proven; on the degenerate sets the proof must refuse rather than be wrong."""
    hi, lo, E, sq, rho = S.split_f16(X)
    approx = S.approx_d2_f16(hi, lo, E, sq, passes=passes, fp32_accumulate=True)
    bound = S.err_bound(X.shape[1], sq, rho, passes)
    ind, proven = S.select_rerank_prove(X, approx, bound)
    ref_ind, ref_dist = O.exact_knn(X, 25)
    exact, tie, bad = O.knn_sets_match(ind[proven], ref_ind[proven], ref_dist[proven])
    assert bad == 0, (name, exact, tie, bad)
    if name == "clusters_d128":
        assert proven.mean() > 0.98
    if name == "tight_blob":
        assert proven.mean() < 0.5   # gaps of ~1e-6 against a bound of ~1e-4: the proof has to give up, and does


@pytest.mark.parametrize("name,X", list(_datasets()), ids=[n for n, _ in _datasets()])
@pytest.mark.parametrize("passes", [1, 2])
def test_packed_key_epilogue_values_within_bound(name, X, passes):
    """The epilogue's packed keys (knn_tc.cu): shifting the column value by 2 max|x|^2 and un-shifting it when a set is flushed
    adds a few roundings at the magnitude of max|x|^2 -- budgeted in err_coef (12 ulps x 4) -- and the shifted value must be
    positive for its bits to order like integers.  Clearing five mantissa bits only LOWERS a stored value (by < 2^-18 of it): the
    proof needs every non-candidate's true approximate distance to be >= the 32nd candidate's stored one, which clearing keeps;
    what a flushed candidate carries must still be a lower bound of its own uncleared value and within bound + 2^-18 of exact."""
    hi, lo, E, sq, rho = S.split_f16(X)
    d = X.shape[1]
    flushed, cleared = S.epilogue_values_packed(hi, lo, E, sq, passes=passes)
    assert (cleared.view(np.uint32) >> 31 == 0).all() and (cleared >= 0).all()       # positive: unsigned order == value order
    exact = S.exact_d2(X)
    bound = S.err_bound(d, sq, rho, passes)[:, None]
    slack = 2.0 ** -18 * 5.0 * float(sq.max())                                        # what clearing five bits can take away
    assert ((flushed.astype(np.float64) - exact) <= bound).all(), name              # never above exact + bound
    assert ((exact - flushed.astype(np.float64)) <= bound + slack).all(), name      # below by at most bound + the cleared bits
    # un-cleared, the shifted arithmetic alone stays inside the bound on both sides
    cs = np.float32(2.0) * np.float32(sq.max())
    ri = np.ldexp(np.float32(1.0), E).astype(np.float32)
    acc = (hi.astype(np.float32) @ hi.astype(np.float32).T + (lo.astype(np.float32) @ hi.astype(np.float32).T if passes == 2 else 0)).astype(np.float32)
    v = ((acc * ri[:, None]).astype(np.float64) * (-2.0 * ri[None, :].astype(np.float64)) + (sq + cs).astype(np.float32)[None, :].astype(np.float64)).astype(np.float32)
    plain = ((v - cs).astype(np.float32) + sq[:, None]).astype(np.float32).astype(np.float64)
    assert (np.abs(plain - exact) <= bound).all(), name


@pytest.mark.parametrize("name,X", list(_hard_datasets()), ids=[n for n, _ in _hard_datasets()])
def test_proven_rows_equal_exact_knn_with_packed_keys(name, X):
    """The same pipeline test as above on what the packed-key epilogue flushes: selection on cleared values (ties by index),
    exact re-rank, proof -- proven rows carry the oracle's lists, on lattice ties and duplicates too."""
    hi, lo, E, sq, rho = S.split_f16(X)
    flushed, _ = S.epilogue_values_packed(hi, lo, E, sq, passes=1)
    bound = S.err_bound(X.shape[1], sq, rho, 1)
    ind, proven = S.select_rerank_prove(X, flushed, bound)
    ref_ind, ref_dist = O.exact_knn(X, 25)
    exact, tie, bad = O.knn_sets_match(ind[proven], ref_ind[proven], ref_dist[proven])
    assert bad == 0, (name, exact, tie, bad)
    if name == "clusters_d128":
        assert proven.mean() > 0.98
