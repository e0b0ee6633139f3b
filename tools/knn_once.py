"""One kNN search at a given shape (for ncu captures of the tensor-core kernel):  python tools/knn_once.py n d [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from graphlearninglayer_b200 import _lib  # noqa: E402

lib = _lib.lib
n, d = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator().manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).cuda()
idx = torch.empty((n, 25), dtype=torch.int32, device="cuda")
dist = torch.empty((n, 25), device="cuda")
info = torch.zeros(_lib.INFO_WORDS, dtype=torch.int32, device="cuda")
wsb = lib.gll_knn_workspace_bytes(n, d, 25)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    _lib.check(lib.gll_knn(X.data_ptr(), n, d, 25, idx.data_ptr(), dist.data_ptr(), info.data_ptr(), ws.data_ptr(), wsb, s), "knn")
torch.cuda.synchronize()
print("ok", int(info[_lib.INFO_KNN_FALLBACK_ROWS]))
