"""Debug tool: per-phase timeline of the on-chip CG kernel on the C4-size system (run on the GPU box)."""
import os, sys
os.environ.setdefault("GLL_B200_CG_PATH", "resident")  # the trace instruments cg_resident.cu, not the one-CTA cg_small.cu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200 import _lib
from oracle.gll_oracle import synth_inputs

shape = sys.argv[1] if len(sys.argv) > 1 else 'c4'
k_lab, m, d, l = (2048, 14336, 512, 10) if shape == 'c4' else (10000, 512, 512, 10)
X, Y, _, yq = synth_inputs(2000, k_lab, m, d, l, 4.5)
Xd = torch.as_tensor(X).cuda(); Yd = torch.as_tensor(Y).cuda()
for _ in range(2): pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
G = 148 if shape == 'c4' else 1
trace = torch.zeros(G * 16 * 16, dtype=torch.int64, device="cuda")
_lib.lib.gll_debug_cg_trace(trace.data_ptr())
pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
torch.cuda.synchronize()
_lib.lib.gll_debug_cg_trace(None)
t = trace.cpu().numpy().reshape(G, 16, 16).astype(np.float64)
t0 = t[:, 0, 0]  # per-CTA origin: clock64 is a per-SM counter
t = t - t0[:, None, None]
t0 = 0.0
names = ["loop top (own update done)", "after barrier B (u visible)", "own gathers done", "all gathers done", "own row sums done",
         "w complete", "own dot products done", "after barrier A", "own column sums done", "", "", "sums known", "", "",
         "scalars done", "own vector update done"]
if shape != 'c4':
    raise SystemExit("the phase timeline instruments the multi-CTA kernel (c4)")
for p in range(2, 5):
    print("pass", p)
    prev = None
    for ph in (0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 14, 15):
        v = (t[:, p, ph] - t0) / 1965.0  # cycles -> us at 1965 MHz
        med = float(np.median(v))
        print(f"   {names[ph]:30s} min {v.min():8.2f} us  median {med:8.2f}  max {v.max():8.2f}   (+{0.0 if prev is None else med - prev:5.2f})")
        prev = med
print(pkg.last_info())
