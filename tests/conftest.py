import glob
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    """Layer fixtures (oracle/make_golden.py); eval_* fixtures (oracle/make_golden_eval.py) have their own loaders."""
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith("eval_"))


def load_golden(name):
    """Returns (fixture dict, X, Y, y_query) with the inputs regenerated from the seed and
    checked against the SHA-256 stored when the reference was run."""
    from oracle.gll_oracle import synth_inputs

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    seed, k_lab, m, d, l = (int(v) for v in g["params"])
    X, Y, _, yq = synth_inputs(seed, k_lab, m, d, l, float(g["sigma"]))
    assert hashlib.sha256(X.tobytes()).hexdigest() == str(g["x_sha256"]), "synthetic input drifted"
    if str(g["label_dtype"]) == "int64":
        Y = Y.astype(np.int64)
    eps = str(g["epsilon"])
    g["eps_arg"] = "auto" if eps == "auto" else float(eps)
    g["tau_arg"] = float(g["tau"])
    return g, X, Y, yq


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
