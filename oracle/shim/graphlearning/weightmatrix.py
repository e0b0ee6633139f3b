"""Exact brute-force ``knnsearch`` (harness; independent of oracle/gll_oracle.exact_knn so the
fixture check is not circular in code): torch.cdist in fp64 without the matmul shortcut,
self forced into slot 0 with distance 0, ties by lower index, distances rounded to fp32
(annoy returns fp32)."""
import numpy as np
import torch


def knnsearch(X, k, method=None, similarity="euclidean", dataset=None, metric="raw", **kw):
    if similarity != "euclidean":
        raise NotImplementedError(similarity)
    Xt = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).double()
    n = Xt.shape[0]
    ind = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    blk = max(1, min(n, int(2e8 // max(1, n))))
    for s in range(0, n, blk):
        e = min(n, s + blk)
        D = torch.cdist(Xt[s:e], Xt, compute_mode="donot_use_mm_for_euclid_dist")
        D[torch.arange(e - s), torch.arange(s, e)] = -1.0
        # stable sort => ties resolved by lower column index
        vals, idx = torch.sort(D, dim=1, stable=True)
        vals, idx = vals[:, :k].clone(), idx[:, :k]
        vals[:, 0] = 0.0
        ind[s:e] = idx.numpy()
        dist[s:e] = vals.float().double().numpy()
    return ind, dist
