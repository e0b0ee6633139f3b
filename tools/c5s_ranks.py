"""Per-call times of the sharded 131072-node graph (c5s) on REAL ranks (torchrun), both CG partitions: one event pair per call."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphlearninglayer_b200 import sharded as sh  # noqa: E402
from graphlearninglayer_b200.losses import custom_ce_loss  # noqa: E402
from graphlearninglayer_b200.synth import synth_inputs  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k_lab, m, d, l = 8192, 122880, 256, 100
X, Y, _, yq = synth_inputs(1000, k_lab, m, d, l, 3.0)
Xd = torch.as_tensor(X).cuda().requires_grad_(True)
Yd = torch.as_tensor(Y).cuda()
yq_d = torch.as_tensor(yq).cuda()
for part in ("columns", "rows-p2p", "columns"):
    def call():
        Xd.grad = None
        pred = sh.ShardedLaplaceLearning.apply(Xd, Yd, 0.0, "auto", None, 0, part)
        custom_ce_loss(pred, yq_d).backward()
    times = []
    for i in range(6):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        times.append(round(e0.elapsed_time(e1), 2))
    if rank == 0:
        print(part, "ms per call:", times, sh.last_info().get("cg_solve_ms"), flush=True)
dist.destroy_process_group()
