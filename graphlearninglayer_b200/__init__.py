"""graphlearninglayer_b200 -- B200 (sm_100a) implementation of the GraphLearningLayer hot path.

    from graphlearninglayer_b200 import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad

The names resolve lazily so that ``python -m graphlearninglayer_b200.build`` can run before the shared library
exists; the first access loads libgll_b200.so through ctypes and fails loudly if it is missing or stale.  There is no
CPU or PyTorch fallback.
"""
__all__ = ["LaplaceLearningSparseHard", "LaplaceLearningSparseHardNormalized", "knn_sym_dist", "stable_conjgrad", "last_info", "GLL",
           "GraphedStep", "BaseSetEvaluator"]
_ELSEWHERE = {"GraphedStep": ".graphed", "BaseSetEvaluator": ".evalcache"}


def __getattr__(name):
    if name in _ELSEWHERE:
        import importlib

        return getattr(importlib.import_module(_ELSEWHERE[name], __name__), name)
    if name in __all__:
        import importlib

        mod = importlib.import_module(".GLL", __name__)
        return mod if name == "GLL" else getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
