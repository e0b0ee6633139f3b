"""One large-graph-mode kNN search (a 131072-row block of a 1M-node, d=256 graph) for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
from graphlearninglayer_b200 import _lib
lib = _lib.lib
n, d, rows = 1 << 20, 256, 131072
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device="cuda"), dim=1)
idx = torch.empty((n, 25), dtype=torch.int32, device="cuda"); dist = torch.empty((n, 25), device="cuda")
info = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32, device="cuda")
wsb = lib.gll_knn_rows_workspace_bytes(n, d, 25, 0, rows); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.gll_knn_rows(X.data_ptr(), n, d, 25, 0, rows, idx.data_ptr(), dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "knn_rows")
torch.cuda.synchronize()
print("ok")
