"""Generate tests/golden/eval_k50.npz: the reference's evaluation routine utils.laplace (utils.py:570-593) run on the
UNMODIFIED /root/reference/GLL.py functions knn_sym_dist (k = 50) and stable_conjgrad (tol 1e-10).

HARNESS ONLY -- build container only.  utils.py itself cannot be imported here (matplotlib, umap, torchvision data
loaders), so its dozen lines are restated below around the two reference functions; graphlearning is the exact-kNN stand-in
of oracle/shim, as for the other fixtures.

    python oracle/make_golden_eval.py
"""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import GLL as REF  # noqa: E402
from graphlearninglayer_b200.synth import synth_inputs  # noqa: E402

assert os.path.realpath(REF.__file__) == "/root/reference/GLL.py", REF.__file__

SEED, K_LAB, M, D, L, SIGMA, KNN, TAU = 23, 300, 1200, 64, 10, 2.0, 50, 1e-8


def main():
    X, Y, y_base, _ = synth_inputs(SEED, K_LAB, M, D, L, SIGMA)
    # ---- utils.py:570-593 ----
    W = REF.knn_sym_dist(X, k=KNN, epsilon="auto")[0]
    Lap = sparse.csgraph.laplacian(W).tocsr()
    label_matrix = Y.astype(np.float64)                       # one_hot_encode(train_labels, n_classes)
    k = label_matrix.shape[0]
    Luu = Lap[k:, k:]
    Lul = Lap[k:, :k]
    m = Luu.shape[0]
    Luu = Luu + sparse.spdiags(TAU * np.ones(m), 0, m, m).tocsr()
    Mv = Luu.diagonal()
    Ms = sparse.spdiags(1 / np.sqrt(Mv + 1e-10), 0, m, m).tocsr()
    Pred = REF.stable_conjgrad(Ms * Luu * Ms, -Ms * Lul @ label_matrix)
    Pred = Ms * Pred
    # ----
    Wc = sparse.csr_matrix(W)
    Wc.sort_indices()
    path = os.path.join(ROOT, "tests", "golden", "eval_k50.npz")
    np.savez_compressed(path, params=np.array([SEED, K_LAB, M, D, L, KNN], dtype=np.int64), sigma=np.float64(SIGMA),
                        tau=np.float64(TAU), x_sha256=np.array(hashlib.sha256(X.tobytes()).hexdigest()),
                        w_indptr=Wc.indptr.astype(np.int32), w_indices=Wc.indices.astype(np.int32), w_data=Wc.data.astype(np.float64),
                        pred=np.asarray(Pred, dtype=np.float64))
    acc = float((np.asarray(Pred).argmax(1) == synth_inputs(SEED, K_LAB, M, D, L, SIGMA)[3]).mean())
    print(f"eval_k50: n={K_LAB + M} nnz(W)={Wc.nnz} accuracy={acc:.3f} -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
