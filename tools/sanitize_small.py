"""Small end-to-end calls for compute-sanitizer (memcheck): every kernel family once, tiny shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200.sharded import ShardedLaplaceLearning
from oracle.gll_oracle import synth_inputs

def run(layer, seed, k_lab, m, d, l, eps, tau, extra=()):
    X, Y, _, yq = synth_inputs(seed, k_lab, m, d, l, 2.0)
    Xt = torch.as_tensor(X).cuda().requires_grad_(True)
    pred = layer(Xt, torch.as_tensor(Y).cuda(), tau, eps, *extra)
    (pred.sum() * 0.5 + (pred ** 2).sum()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(pred).all() and torch.isfinite(Xt.grad).all()

for path in ("tc", "simt"):
    os.environ["GLL_B200_KNN_PATH"] = path
    run(pkg.LaplaceLearningSparseHard.apply, 0, 150, 390, 72, 7, "auto", 0.0)      # cluster CG (eight CTAs, DSMEM), ragged tiles
    run(pkg.LaplaceLearningSparseHard.apply, 1, 100, 2100, 40, 10, 1.0, 0.07)     # multi-CTA on-chip CG
    run(pkg.LaplaceLearningSparseHard.apply, 4, 300, 1500, 64, 10, "auto", 0.0)    # cluster CG, two items per thread
os.environ["GLL_B200_CG_PATH"] = "small"
run(pkg.LaplaceLearningSparseHard.apply, 0, 150, 390, 72, 7, "auto", 0.0)          # one-CTA CG
del os.environ["GLL_B200_CG_PATH"]
for ares in ("0", "1"):
    os.environ["GLL_B200_KNN_ARES"] = ares
    os.environ["GLL_B200_KNN_PAIR"] = ares
    run(pkg.LaplaceLearningSparseHard.apply, 5, 200, 1300, 96, 6, "auto", 0.0)     # resident A operand / CTA pairs
del os.environ["GLL_B200_KNN_ARES"], os.environ["GLL_B200_KNN_PAIR"]
for path in ("tc",):
    pass
os.environ["GLL_B200_KNN_PATH"] = "tc"
os.environ["GLL_B200_CG_PATH"] = "streaming"
run(pkg.LaplaceLearningSparseHard.apply, 2, 100, 900, 33, 5, "auto", 0.0)          # streaming CG, d % 4 != 0
del os.environ["GLL_B200_CG_PATH"]
run(ShardedLaplaceLearning.apply, 3, 200, 700, 64, 13, "auto", 0.0, (None, 3))     # virtual ranks
print("sanitize_small OK")
