"""Put this directory (and the repo root) on PYTHONPATH ahead of the reference checkout and the reference's
``from GLL import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad`` (utils.py:25, FullySup.py:15,
compare_to_mlp.py:13) resolves to the B200 implementation.  See INTEGRATION.md."""
from graphlearninglayer_b200.GLL import LaplaceLearningSparseHard, knn_sym_dist, stable_conjgrad  # noqa: F401
