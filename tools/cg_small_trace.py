"""Debug tool: per-phase timeline (SM cycles, thread 0) of the one-CTA CG kernel cg_small.cu on the C2 minibatch system."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200 import _lib
from oracle.gll_oracle import synth_inputs

X, Y, _, yq = synth_inputs(1000, 10000, 512, 512, 10, 4.5)
Xd = torch.as_tensor(X).cuda(); Yd = torch.as_tensor(Y).cuda()
for _ in range(2): pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
trace = torch.zeros(16 * 8 + 8 * 8, dtype=torch.int64, device="cuda")
_lib.lib.gll_debug_cg_trace(trace.data_ptr())
pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
torch.cuda.synchronize()
_lib.lib.gll_debug_cg_trace(None)
raw = trace.cpu().numpy()
per_cta = raw[128:].reshape(8, 8)
t = raw[:128].reshape(16, 8).astype(np.float64)
t = (t - t[0, 0]) / 1965.0
names = ["loop top", "A done+sync", "B sums done", "C scalars done", "A spmv done(t0)", "A stores done(t0)", "A stores done(t511)", "A row range read(t0)"]
if os.environ.get("GLL_B200_CG_PATH", "") == "":  # cluster kernel (cg_cluster.cu): its own phase names
    names = ["loop top", "after cluster sync 1", "sums known", "scalars + flag", "A done (products stored)", "B done (partials sent)", "D done", "-"]
print(f"kernel entry -> first loop top: {-t[15, 7]:.2f} us")
for p in range(8):
    print("pass", p, "  ".join(f"{names[ph]}: {t[p, ph]:.2f}" for ph in ((0, 4, 5, 1, 2, 3, 6) if names[7] == "-" else (0, 7, 4, 5, 6, 1, 2, 3)) if t[p, ph] > -1e6))
print(pkg.last_info())

if per_cta.any():
    for c in range(8):
        a, b, e, mxl, nz = per_cta[c][:5]
        print(f"CTA {c}: A (all threads) {(b - a) / 1965.0:.2f} us, B {(e - b) / 1965.0:.2f} us, longest row {mxl} edges, {nz} edges in the slice")
