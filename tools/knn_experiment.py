"""Timing experiments on the tensor-core kNN kernel (GLL_B200_KNN_DEBUG / GLL_B200_KNN_PAIR knobs): per-kernel ms of K1 at a
given shape.  Results of debug modes are wrong by construction; this only reads the clock.
    python tools/knn_experiment.py [n d]"""
import os
import subprocess
import sys

CODE = r'''
import sys, torch, numpy as np
sys.path.insert(0, ".")
from graphlearninglayer_b200 import _lib
lib = _lib.lib
n, d = int(sys.argv[1]), int(sys.argv[2])
from graphlearninglayer_b200.synth import synth_inputs
X = torch.as_tensor(synth_inputs(1000, n - 512, 512, d, 10, 4.5)[0]).cuda()  # the benchmark's Gaussian clusters
idx = torch.empty((n, 25), dtype=torch.int32, device="cuda"); dist = torch.empty((n, 25), device="cuda")
info = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32, device="cuda")
wsb = lib.gll_knn_workspace_bytes(n, d, 25); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.gll_knn(X.data_ptr(), n, d, 25, idx.data_ptr(), dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "knn")
for _ in range(3): run()
torch.cuda.synchronize()
lib.gll_profile_enable(1); _lib.profile_collect()
for _ in range(10): run()
torch.cuda.synchronize()
p = _lib.profile_collect()
print({k: round(v[0] / v[1], 4) for k, v in p.items()}, "fallback_rows", int(info[_lib.INFO_KNN_FALLBACK_ROWS].item()))
'''

n, d = (sys.argv[1:3] + ["10512", "512"])[:2] if len(sys.argv) >= 3 else ("10512", "512")
F2 = {"GLL_B200_KNN_SPLIT": "f16x2"}
PAIR = {"GLL_B200_KNN_PAIR": "1"}
# one pass (default) / with CTA pairs / two passes, each also without insertions (DEBUG=1) and as MMA + TMA pipeline alone (2)
ENVS = ({}, {"GLL_B200_KNN_DEBUG": "1"}, {"GLL_B200_KNN_DEBUG": "2"}, PAIR, dict(PAIR, GLL_B200_KNN_DEBUG="1"),
        dict(PAIR, GLL_B200_KNN_DEBUG="2"), F2, dict(F2, GLL_B200_KNN_DEBUG="2"), dict(F2, **PAIR))
if os.environ.get("KNN_EXPERIMENT_ENVS") == "default":  # the default configuration, and without the set insertions
    ENVS = ({}, {"GLL_B200_KNN_DEBUG": "1"})
for env in ENVS:
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, "-c", CODE, n, d], env=e, capture_output=True, text=True)
    print(env, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
