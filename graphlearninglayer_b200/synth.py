"""Synthetic workload generator: Gaussian class clusters, rows L2-normalised, labeled ("base") rows
first.  numpy only; shared by bench.py, the tools and (re-exported) the test oracle, so that every arm sees the same inputs."""
from __future__ import annotations

import numpy as np


def synth_inputs(seed: int, k_lab: int, m: int, d: int, l: int, sigma: float):
    """Returns X (n,d) float32, Y (k_lab,l) float32 one-hot, y_base (k_lab,), y_query (m,)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((l, d))
    y_base = np.arange(k_lab) % l
    y_query = rng.integers(0, l, size=m)
    y = np.concatenate([y_base, y_query])
    X = centres[y] + sigma * rng.standard_normal((k_lab + m, d))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = X.astype(np.float32)
    Y = np.zeros((k_lab, l), dtype=np.float32)
    Y[np.arange(k_lab), y_base] = 1.0
    return X, Y, y_base, y_query
