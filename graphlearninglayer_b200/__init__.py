"""graphlearninglayer_b200 -- B200 (sm_100a) implementation of the GraphLearningLayer hot path.

    from graphlearninglayer_b200 import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad

Importing the package loads libgll_b200.so (built in-tree by ``python -m graphlearninglayer_b200.build``) and fails
loudly if it is missing; there is no CPU or PyTorch fallback.
"""
from .GLL import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad, last_info  # noqa: F401

__all__ = ["LaplaceLearningSparseHard", "knn_sym_dist", "stable_conjgrad", "last_info"]
