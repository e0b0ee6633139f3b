"""Under torchrun (>= 2 GPUs): the peer-memory row-partitioned CG against the column partition on the same graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle.gll_oracle import synth_inputs, max_rel
from graphlearninglayer_b200.sharded import ShardedLaplaceLearning, last_info

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (k_lab, m, d, l, eps, tau) in [(700, 2300, 96, 13, "auto", 0.0), (1024, 5120, 256, 10, 1.0, 0.07)]:
    X, Y, _, yq = synth_inputs(31, k_lab, m, d, l, 2.5)
    res = {}
    for part in ("columns", "rows", "rows-p2p", "rows-p2p"):   # twice: the cached symmetric block is reused with new epochs
        Xt = torch.as_tensor(X).cuda().requires_grad_(True)
        pred = ShardedLaplaceLearning.apply(Xt, torch.as_tensor(Y).cuda(), tau, eps, None, 0, part)
        tgt = torch.nn.functional.one_hot(torch.as_tensor(yq).cuda(), l).to(pred.dtype)
        (-torch.sum(tgt * torch.log(pred + 1e-8)) / m).backward()
        torch.cuda.synchronize()
        info = last_info()
        res.setdefault(part, []).append((pred.detach().cpu().numpy(), Xt.grad.cpu().numpy(), info["cg_iters_fwd"], info["cg_iters_bwd"], info["status"]))
    ref = res["columns"][0]
    for part, runs in res.items():
        for p, g, itf, itb, st in runs:
            e1, e2 = max_rel(p, ref[0]), max_rel(g, ref[1])
            good = e1 < 2e-6 and e2 < 2e-6 and (st & ~8) == 0
            ok &= good
            if rank == 0:
                print(f"n={k_lab + m} l={l} eps={eps}: {part:9s} iters {itf}/{itb} pred err {e1:.2e} dX err {e2:.2e} {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("P2P CHECK", "PASSED" if t.item() else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
