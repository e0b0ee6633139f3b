"""Phase timeline of the fused graph + weights kernel (graph.cu) on the C2 graph: %globaltimer of CTA 0 at the end of each phase."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200 import _lib
from oracle.gll_oracle import synth_inputs

shape = (1000, 10000, 512, 512, 10, 4.5) if (len(sys.argv) < 2 or sys.argv[1] == "c2") else (2, 2048, 14336, 512, 10, 4.5)
X, Y, _, yq = synth_inputs(*shape)
Xd = torch.as_tensor(X).cuda(); Yd = torch.as_tensor(Y).cuda()
for _ in range(2): pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
trace = torch.zeros(16 * 8 + 8 * 8, dtype=torch.int64, device="cuda")
_lib.lib.gll_debug_cg_trace(trace.data_ptr())
import os as _os
_os.environ["GLL_B200_CG_PATH"] = "streaming"  # the CG kernels write their own timelines into the same buffer: keep them out
pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
torch.cuda.synchronize()
_lib.lib.gll_debug_cg_trace(None)
raw = trace.cpu().numpy()
print("raw stamps", raw[64:71])
t = raw[64:71].astype(np.float64)
names = ["0 count (reverse-edge test)", "1 scan row_ptr", "2 fill", "3 sort + weights", "4 scan uu_ptr", "5 L_uu fill (to kernel end)"]
for i, nm in enumerate(names):
    print(f"phase {nm:32s} {(t[i + 1] - t[i]) / 1e3:7.2f} us")
print(f"total {(t[6] - t[0]) / 1e3:.2f} us", pkg.last_info())
