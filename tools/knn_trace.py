"""SM-clock timeline of CTA 0 of the tensor-core kNN Gram kernel (gll_debug_knn_trace): where a unit's time goes in the MMA
warp, two epilogue warps (the two column halves of one row quarter) and the A-tile TMA producer.

    python tools/knn_trace.py c2            # 10000 + 512 nodes, d = 512 (units dealt to CTAs)
    python tools/knn_trace.py rows          # one 131072-row block of a 1M-node graph, d = 256 (whole row tiles per CTA)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from graphlearninglayer_b200 import _lib  # noqa: E402

lib = _lib.lib
mode = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = torch.Generator(device="cuda").manual_seed(0)
if mode == "rows":
    n, d, rows = 1 << 20, 256, 131072
    cen = torch.randn(100, d, generator=g, device="cuda")
    X = torch.nn.functional.normalize(cen[torch.arange(n, device="cuda") % 100] + 3.0 * torch.randn(n, d, generator=g, device="cuda"), dim=1)
else:
    n, d, rows = 10512, 512, 10512
    cen = torch.randn(10, d, generator=g, device="cuda")
    X = torch.nn.functional.normalize(cen[torch.arange(n, device="cuda") % 10] + 3.0 * torch.randn(n, d, generator=g, device="cuda"), dim=1)
idx = torch.empty((n, 25), dtype=torch.int32, device="cuda")
dist = torch.empty((n, 25), device="cuda")
info = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32, device="cuda")
wsb = lib.gll_knn_rows_workspace_bytes(n, d, 25, 0, rows)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.gll_knn_rows(X.data_ptr(), n, d, 25, 0, rows, idx.data_ptr(), dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "knn_rows")


run()
torch.cuda.synchronize()
W, U, P = 4, 1024, 8
trace = torch.zeros(W * U * P, dtype=torch.int64, device="cuda")
lib.gll_debug_knn_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
lib.gll_debug_knn_trace(None)
t = trace.cpu().numpy().reshape(W, U, P)
nu = int((t[1, :, 0] != 0).sum())
print(f"mode {mode}: n={n} d={d}; CTA 0 traced {nu} units")
if nu < 4:
    sys.exit(0)
t0 = t[1, 0, 0]


def col(w, ph):
    return (t[w, :nu, ph] - t0).astype(np.float64)


def report(lo, hi, label):
    hi = min(hi, nu)
    if hi - lo < 2:
        return
    sl = slice(lo, hi)
    per_unit = (col(1, 0)[hi - 1] - col(1, 0)[lo]) / (hi - 1 - lo)
    print(f"-- units [{lo}, {hi}) {label}: {per_unit:.0f} cycles per unit")
    for w, name in ((1, "epilogue warp 2 (half 0)"), (2, "epilogue warp 6 (half 1)")):
        a = [col(w, p)[sl] for p in range(5)]
        print(f"   {name}: barriers+norms {np.mean(a[1] - a[0]):.0f}  threshold load {np.mean(a[2] - a[1]):.0f}  wait accumulator {np.mean(a[3] - a[2]):.0f}"
              f"  chunks {np.mean(a[4] - a[3]):.0f}  rounds/unit {np.mean(t[w, sl, 5]):.2f}  end->next top {np.mean(col(w, 0)[lo + 1:hi] - a[4][:-1]):.0f}")
    m = [col(0, p)[sl] for p in range(4)]
    print(f"   MMA warp: wait for a drained accumulator {np.mean(m[1] - m[0]):.0f}  wait first operands {np.mean(m[2] - m[1]):.0f}  first->last K block {np.mean(m[3] - m[2]):.0f}"
          f"  unit to unit {np.mean(np.diff(m[0])):.0f}")
    mm = [col(0, p)[sl] for p in range(8)]
    print(f"             operands of K block 0 / 1 / 2 / last ready at +{np.mean(mm[2] - mm[1]):.0f} / +{np.mean(mm[4] - mm[1]):.0f} / +{np.mean(mm[5] - mm[1]):.0f} / "
          f"+{np.mean(mm[3] - mm[1]):.0f}; issue of K block 0 took {np.mean(mm[6] - mm[2]):.0f}, of the last {np.mean(mm[7] - mm[3]):.0f}; "
          f"last issue -> next unit {np.mean(mm[0][1:] - mm[7][:-1]):.0f}")
    pa = [col(3, p)[sl] for p in range(2)]
    print(f"   producer A: unit start -> last stage free {np.mean(pa[1] - pa[0]):.0f}  unit to unit {np.mean(np.diff(pa[0])):.0f}")


report(0, 4, "(start)")
report(4, 16, "")
report(16, 64, "")
report(64, 256, "")
report(256, 1024, "")

blocks = int((t[1, :, 7] != 0).sum())
if blocks >= 2:
    cyc = np.diff(t[1, :blocks, 7].astype(np.float64)) / 1024.0
    ns = np.diff(t[1, :blocks, 6].astype(np.float64)) / 1024.0
    print("-- whole CTA, per block of 1024 units: cycles per unit", [int(c) for c in cyc])
    print("                                        ns per unit    ", [int(x) for x in ns])
    print("                                        SM clock, GHz  ", [round(c / x, 2) for c, x in zip(cyc, ns)])
