"""cProfile of the host side of fwd+bwd steps (where does the Python / launch time go?)."""
import cProfile, pstats, sys, time, io
sys.path.insert(0, ".")
import torch
import graphlearninglayer_b200 as pkg
from graphlearninglayer_b200.losses import custom_ce_loss
from oracle.gll_oracle import synth_inputs

k_lab, m, d, l = 4096, 512, 512, 10
X, Y, _, yq = synth_inputs(1000, k_lab, m, d, l, 4.5)
Xd = torch.as_tensor(X).cuda().requires_grad_(True); Yd = torch.as_tensor(Y).cuda(); yd = torch.as_tensor(yq).cuda()
def step():
    Xd.grad = None
    pred = pkg.LaplaceLearningSparseHard.apply(Xd, Yd, 0.0, "auto")
    loss = custom_ce_loss(pred, yd)
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / N:.3f} ms/step, wall incl. drain {1e3 * (t2 - t0) / N:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
