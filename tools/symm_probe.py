"""Probe: does torch.distributed._symmetric_memory give peer pointers over NVLink in this image? (run under torchrun)"""
import os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty((1 << 20,), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    print(rank, "symm ok: ptrs", [hex(p) for p in ptrs], "signal pads", [hex(int(p)) for p in hdl.signal_pad_ptrs], flush=True)
    # write my rank into every peer's buffer at slot [rank] through the peer views
    for p in range(world):
        peer = hdl.get_buffer(p, (1 << 20,), torch.float32)
        peer[rank] = float(rank + 1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(rank, "local view after peers wrote:", t[:world].tolist(), flush=True)
    # peer copy bandwidth
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    big = symm_mem.empty((64 << 20,), dtype=torch.float32, device=dev)
    h2 = symm_mem.rendezvous(big, dist.group.WORLD)
    pb = h2.get_buffer((rank + 1) % world, (64 << 20,), torch.float32)
    src = torch.ones(64 << 20, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pb.copy_(src)
    e1.record(); torch.cuda.synchronize()
    print(rank, "peer store bandwidth GB/s:", 5 * 256e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm FAILED:", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
