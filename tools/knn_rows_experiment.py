"""Timing experiments on the tensor-core kNN kernel in its large-graph mode (whole row tiles per CTA): one row block of a
1M-node, d=256 graph through gll_knn_rows, under the GLL_B200_KNN_DEBUG / GLL_B200_KNN_SHARE knobs."""
import os
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from graphlearninglayer_b200 import _lib
lib = _lib.lib
n, d, rows = 1 << 20, 256, 131072
g = torch.Generator(device="cuda").manual_seed(0)   # 100 Gaussian clusters, sigma 3 (the benchmark's distribution), made on the device
cen = torch.randn(100, d, generator=g, device="cuda")
X = torch.nn.functional.normalize(cen[torch.arange(n, device="cuda") % 100] + 3.0 * torch.randn(n, d, generator=g, device="cuda"), dim=1)
idx = torch.empty((n, 25), dtype=torch.int32, device="cuda"); dist = torch.empty((n, 25), device="cuda")
info = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32, device="cuda")
wsb = lib.gll_knn_rows_workspace_bytes(n, d, 25, 0, rows); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.gll_knn_rows(X.data_ptr(), n, d, 25, 0, rows, idx.data_ptr(), dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "knn_rows")
run(); torch.cuda.synchronize()
lib.gll_profile_enable(1); _lib.profile_collect()
for _ in range(3): run()
torch.cuda.synchronize()
p = _lib.profile_collect()
ms = p["knn_gram_topk_tcgen05"][0] / p["knn_gram_topk_tcgen05"][1]
import os
npass = 2 if os.environ.get("GLL_B200_KNN_SPLIT") == "f16x2" else 1
print({k: round(v[0] / v[1], 3) for k, v in p.items()}, "issued PFLOP/s", round(npass * 2.0 * rows * n * d / ms / 1e12, 3),
      "fallback_rows", int(info[_lib.INFO_KNN_FALLBACK_ROWS].item()))
'''
ENVS = ({}, {"GLL_B200_KNN_DEBUG": "1"}, {"GLL_B200_KNN_DEBUG": "2"}, {"GLL_B200_KNN_SPLIT": "f16x2"},
        {"GLL_B200_KNN_SPLIT": "f16x2", "GLL_B200_KNN_DEBUG": "2"})
if os.environ.get("KNN_EXPERIMENT_ENVS") == "pair":  # CTA pairs (default here) against single CTAs, each also as the pipeline alone
    ENVS = ({}, {"GLL_B200_KNN_DEBUG": "1"}, {"GLL_B200_KNN_DEBUG": "2"}, {"GLL_B200_KNN_PAIR": "0"},
            {"GLL_B200_KNN_PAIR": "0", "GLL_B200_KNN_DEBUG": "2"})
if os.environ.get("KNN_EXPERIMENT_ENVS") == "default":  # the default configuration, and without the set insertions
    ENVS = ({}, {"GLL_B200_KNN_DEBUG": "1"})
for env in ENVS:
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
    print(env, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:], flush=True)
