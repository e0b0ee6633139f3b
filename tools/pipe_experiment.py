"""HostPipeline depth sweep at the C2 shape: device-timed ms per call and host-side ms per submit."""
import sys, time
sys.path.insert(0, ".")
import torch
from oracle.gll_oracle import synth_inputs
from graphlearninglayer_b200.hostpipe import HostPipeline

k_lab, m, d, l = 10000, 512, 512, 10
X, Y, _, yq = synth_inputs(1000, k_lab, m, d, l, 4.5)
Xh, Yh = torch.as_tensor(X).pin_memory(), torch.as_tensor(Y).pin_memory()
tgt = torch.nn.functional.one_hot(torch.as_tensor(yq), l).to(torch.float64).cuda()
for depth in (1, 2, 3, 4):
    pipe = HostPipeline(k_lab + m, d, k_lab, l, "cuda", loss_fn=lambda p, s: -torch.sum(tgt * torch.log(p + 1e-8)) / m, depth=depth)
    for _ in range(5):
        if pipe.outstanding == depth: pipe.collect()
        pipe.submit(Xh, Yh)
    pipe.drain(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); host = 0.0; N = 40
    for _ in range(N):
        if pipe.outstanding == depth: pipe.collect()
        t = time.perf_counter(); pipe.submit(Xh, Yh); host += time.perf_counter() - t
    pipe.drain(); e1.record(); torch.cuda.synchronize()
    print(f"depth {depth}: {e0.elapsed_time(e1) / N:.3f} ms/call device-timed, host submit {1e3 * host / N:.3f} ms/call", flush=True)
