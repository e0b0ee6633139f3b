"""Evaluation loop with a fixed base set: per-call kernel times of BaseSetEvaluator against the plain layer
forward at the C2 shape (10000 base + 512 batch, d = 512).  Run on the GPU box: python tools/eval_cache_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200 import _lib
from graphlearninglayer_b200.evalcache import BaseSetEvaluator
from graphlearninglayer_b200.synth import synth_inputs

X, Y, _, yq = synth_inputs(1000, 10000, 512 * 4, 512, 10, 4.5)
base, Yt = torch.as_tensor(X[:10000]).cuda(), torch.as_tensor(Y).cuda()
batches = [torch.as_tensor(X[10000 + 512 * i:10000 + 512 * (i + 1)]).cuda() for i in range(4)]
ev = BaseSetEvaluator(base, Yt, 0.0, "auto")


def timed(fn, reps=10):
    for b in batches: fn(b)
    torch.cuda.synchronize()
    _lib.lib.gll_profile_enable(1); _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for b in batches: fn(b)
    e1.record(); torch.cuda.synchronize()
    prof = _lib.profile_collect(); _lib.lib.gll_profile_enable(0)
    calls = reps * len(batches)
    return e0.elapsed_time(e1) / calls, {k: round(v[0] / calls, 4) for k, v in prof.items()}


def plain(b):
    with torch.no_grad():
        return pkg.LaplaceLearningSparseHard.apply(torch.cat((base, b), 0), Yt, 0.0, "auto")


ms_c, k_c = timed(ev)
ms_p, k_p = timed(plain)
same = all(torch.equal(ev(b), plain(b)) for b in batches)
print(f"plain forward  : {ms_p:.4f} ms per call (profiled), kernel ms per call {k_p}")
print(f"base-set cache : {ms_c:.4f} ms per call (profiled), kernel ms per call {k_c}")
print(f"K1 (Gram + top-k) {k_p.get('knn_gram_topk_tcgen05', 0) / max(k_c.get('knn_gram_topk_tcgen05', 1e-9), 1e-9):.2f}x faster, call {ms_p / ms_c:.2f}x, bit-identical predictions: {same}")
