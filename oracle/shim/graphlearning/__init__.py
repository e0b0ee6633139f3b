"""Stand-in for the ``graphlearning`` package (absent from this image, no network) so that the
UNMODIFIED reference ``GLL.py`` can be imported for fixture generation (oracle/make_golden.py).

HARNESS ONLY.  Implements just the two call sites of the hot path:
``gl.weightmatrix.knnsearch`` (GLL.py:183) -- here EXACT brute force instead of annoy -- and
``gl.graph(W).gradient(u)`` (GLL.py:111-119): sparse matrix (u_j - u_i) on the pattern of W.
"""
import numpy as np
import scipy.sparse as sparse

from . import weightmatrix  # noqa: F401


class graph:
    def __init__(self, W):
        self.weight_matrix = sparse.csr_matrix(W)
        self.num_nodes = self.weight_matrix.shape[0]

    def gradient(self, u, weighted=False, p=0.0):
        P = self.weight_matrix.copy()
        P.data = np.ones_like(P.data)
        n = self.num_nodes
        U = sparse.spdiags(np.asarray(u, dtype=np.float64), 0, n, n)
        return (P @ U - U @ P).tocsr()
