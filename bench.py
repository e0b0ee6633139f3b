#!/usr/bin/env python
"""bench.py -- GLL hot-path benchmark (BASELINE.json metric: GLL fwd+bwd calls/s, CG iterations/s vs HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl b200|reference]

One "step" = one call of the hot path on one synthetic graph:
    pred = LaplaceLearningSparseHard.apply(X, Y, tau, eps); loss = custom_ce_loss(pred, y); loss.backward()
Workloads (BASELINE.json configs): c2 = 10000 base + 512 batch, d=512, eps='auto' (default: the
config the metric is quoted on for one GPU); c3 = 4096 + 512; c4 = 2048 + 14336 (CG roofline study).
N > 1: every rank runs its own independent graph of the same shape (different seed) -- the layer has no
parameters, so there is no data-path collective (weak scaling).

value  : calls/s with X, Y resident in HBM, timed per step with CUDA events on the launch stream, L2 flushed
         between steps, max over ranks.
e2e    : the same call through the public autograd API starting from pinned HOST buffers: H2D of X and Y, the call,
         D2H of pred and dX, all inside the timed region.
roofline / kernels : a second pass of the same K steps with the library's per-kernel CUDA-event brackets on
         (gll_profile_enable), so the timed `value` pass carries no instrumentation.
cpu_baseline : the fp64 numpy/scipy oracle (a port of the reference's GLL.py, see oracle/gll_oracle.py) on this
         box's host cores, rank 0 at N=1 only, bounded sample.
--impl reference : only the oracle port, all host threads, same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name: (k_lab, m, d, l, sigma, tau, eps)
WORKLOADS = {
    "c2": (10000, 512, 512, 10, 4.5, 0.0, "auto"),
    "c3": (4096, 512, 512, 10, 4.5, 0.0, "auto"),
    "c4": (2048, 14336, 512, 10, 4.5, 0.0, "auto"),
    "c5": (65536, 983040, 256, 100, 3.0, 0.0, "auto"),   # ONE graph sharded over the ranks (strong scaling)
    "c5s": (8192, 122880, 256, 100, 3.0, 0.0, "auto"),   # 1/8-size version of c5 for quick runs
}
SHARDED = ("c5", "c5s")
WORKLOAD_DESC = {
    "c2": "GLL classifier head, 10000 base + 512 batch, d=512, l=10, eps='auto', tau=0, k=25",
    "c3": "data-parallel GLL step, 4096 base + 512 batch per rank, d=512, l=10, eps='auto', tau=0, k=25",
    "c4": "large single graph, 2048 base + 14336 unlabeled, d=512, l=10, eps='auto', tau=0, k=25, CG tol 1e-7",
    "c5": "sharded Laplace learning, ONE graph of n=2^20 nodes (65536 labeled), d=256, l=100, eps='auto', tau=0, k=25; "
          "kNN/backward row-sharded, CG column-sharded, NCCL all-gathers",
    "c5s": "sharded Laplace learning, ONE graph of n=131072 nodes (8192 labeled), d=256, l=100, eps='auto', tau=0, k=25; "
           "kNN/backward row-sharded, CG column-sharded, NCCL all-gathers",
}
METRIC = "gll_fwd_bwd_calls_per_sec"
UNIT = "calls/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        return dict(hbm_gbs=float(z["hbm_gbs"]), tflops=float(z["bf16_tflops"]), source="measured (MEASURED_PEAKS.json, burst)",
                    tflops_sustained=float(z.get("bf16_tflops_sustained", z["bf16_tflops"])))
    return dict(hbm_gbs=6650.0, tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi sampled every 100 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (checker code; never on the GPU product path)
# ----------------------------------------------------------------------------------------------------------------
def cpu_calls(workload, seed, n_calls, budget_s):
    from oracle import gll_oracle as O

    k_lab, m, d, l, sigma, tau, eps = WORKLOADS[workload]
    X, Y, _, yq = O.synth_inputs(seed, k_lab, m, d, l, sigma)
    times = []
    t_all = time.perf_counter()
    for _ in range(n_calls):
        t0 = time.perf_counter()
        O.fwd_bwd(X, Y, yq, tau, eps)  # exact kNN + scipy graph + LU/CG solves + per-edge backward, fp64
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    return times


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    # warm-up calls are run but bounded to one when a call takes seconds
    warm = cpu_calls(args.workload, 1000, min(args.warmup, 1), 60.0)
    times = cpu_calls(args.workload, 1000, args.steps, 1e9)
    total = float(np.sum(times))
    val = len(times) / total
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": len(warm), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic gaussian clusters (graphlearninglayer_b200.synth), L2-normalised",
        "config": {"workload": args.workload + ": " + WORKLOAD_DESC[args.workload],
                   "what": "oracle port of /root/reference GLL.py (exact kNN instead of annoy, scipy sparse, SuperLU/CG), "
                           "one process on the host cores; the Python reference itself cannot travel to the GPU box"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full fwd+bwd calls of the workload"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def algorithmic_work(name, shp, info):
    """(bound, work per launch, unit) for the kernels with a roofline entry."""
    n, d, l, m, k = shp["n"], shp["d"], shp["l"], shp["m"], 25
    E, Euu = info["nnz"], info["nnz_uu"]
    if name.startswith("knn_gram_topk"):
        return "tensor", 2.0 * n * n * d, "flop"
    if name == "cg_persistent":
        # B_iter = 8 E_uu + 4 (m+1) + 11 * 4 m l per iteration; one launch = one solve; fwd and bwd solves averaged
        iters = 0.5 * (info["cg_iters_fwd"] + info["cg_iters_bwd"])
        return "hbm", iters * (8.0 * Euu + 4.0 * (m + 1) + 44.0 * m * l), "B"
    if name == "row_gather":
        return "hbm", 8.0 * E + 8.0 * n * d + 12.0 * n, "B"
    if name == "edge_grad":
        return "hbm", 16.0 * E + 8.0 * n * l + 8.0 * n, "B"
    if name == "knn_rerank":
        return "hbm", 4.0 * n * d + 8.0 * n * 32 + 8.0 * n * k, "B"
    if name == "edge_weights":
        return "hbm", 8.0 * n * k + 12.0 * E + 4.0 * (n + 1) + 4.0 * m * l, "B"
    return "hbm", None, "B"


def run_b200(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    import graphlearninglayer_b200 as pkg
    from graphlearninglayer_b200 import _lib, ranks
    from graphlearninglayer_b200.synth import synth_inputs  # numpy input generator (the oracle is not imported by this arm)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def step_fn(shape_name, seed):
        k_lab, m, d, l, sigma, tau, eps = WORKLOADS[shape_name]
        if shape_name in SHARDED:
            from graphlearninglayer_b200.sharded import ShardedLaplaceLearning

            layer = ShardedLaplaceLearning.apply  # every rank passes the same graph, collectives inside
        else:
            layer = pkg.LaplaceLearningSparseHard.apply
        X, Y, _, yq = synth_inputs(seed, k_lab, m, d, l, sigma)
        Xh = torch.as_tensor(X).pin_memory()
        Yh = torch.as_tensor(Y).pin_memory()
        Xd = Xh.to(dev).requires_grad_(True)
        Yd = Yh.to(dev)
        yq_d = torch.as_tensor(yq).to(dev)
        if args.loss == "torch":   # the reference function body, op by op (losses.py:128-136)
            tgt = torch.nn.functional.one_hot(yq_d, l).to(torch.float64)

            def ce(pred):
                return -torch.sum(tgt * torch.log(pred + 1e-8)) / m
        else:                      # the same function as one kernel (graphlearninglayer_b200/losses.py)
            from graphlearninglayer_b200.losses import custom_ce_loss

            def ce(pred):
                return custom_ce_loss(pred, yq_d)
        predh = torch.empty((m, l), dtype=torch.float64).pin_memory()
        dXh = torch.empty((k_lab + m, d), dtype=torch.float32).pin_memory()
        Xe = torch.empty_like(Xd).requires_grad_(True)
        Ye = torch.empty_like(Yd)

        split = {"ev": None}  # set to a list to have resident() record (start, after forward+loss, end) events

        def resident():
            Xd.grad = None
            ev = None
            if split["ev"] is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            pred = layer(Xd, Yd, tau, eps)
            loss = ce(pred)  # custom_ce_loss, losses.py:128-136
            if ev is not None:
                ev[1].record()
            loss.backward()
            if ev is not None:
                ev[2].record()
                split["ev"].append(ev)
            return loss

        def e2e():
            Xe.grad = None
            with torch.no_grad():
                Xe.copy_(Xh, non_blocking=True)
                Ye.copy_(Yh, non_blocking=True)
            pred = layer(Xe, Ye, tau, eps)
            loss = ce(pred)
            loss.backward()
            predh.copy_(pred.detach(), non_blocking=True)
            dXh.copy_(Xe.grad, non_blocking=True)
            return loss

        def e2e_pipelined(steps, warmup):
            """Same host-buffer calls, double buffered (graphlearninglayer_b200.hostpipe): H2D of step i+1 and D2H of
            step i-1 overlap the kernels of step i.  Returns the device-timed ms for `steps` calls, all copies inside."""
            from graphlearninglayer_b200.hostpipe import HostPipeline

            pipe = HostPipeline(k_lab + m, d, k_lab, l, dev, tau=tau, epsilon=eps, layer=layer, depth=3,
                                loss_fn=lambda pred, slot: ce(pred))
            for _ in range(warmup):
                if pipe.outstanding == pipe.depth:
                    pipe.collect()
                pipe.submit(Xh, Yh)
            pipe.drain()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(steps):
                if pipe.outstanding == pipe.depth:
                    pipe.collect()
                pipe.submit(Xh, Yh)
            pipe.drain()
            ev1.record()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            ms = ev0.elapsed_time(ev1)
            _, ms, _ = ranks.aggregate_throughput(steps, ms, device=str(dev))
            return ms, pipe.h2d_bytes, pipe.d2h_bytes

        shp = dict(n=k_lab + m, d=d, l=l, m=m, k_lab=k_lab)
        h2d = Xh.numel() * 4 + Yh.numel() * 4
        d2h = predh.numel() * 8 + dXh.numel() * 4
        return resident, e2e, shp, h2d, d2h, e2e_pipelined, split

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.zero_()          # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        tot_ms = sum(a.elapsed_time(b) for a, b in ev)
        # whole-job time = max over ranks (graphlearninglayer_b200.ranks: all_reduce MAX; no data-path collective)
        _, tot_ms, _ = ranks.aggregate_throughput(steps, tot_ms, device=str(dev))
        return tot_ms

    sharded = args.workload in SHARDED
    resident, e2e, shp, h2d, d2h, e2e_pipelined, split = step_fn(args.workload, 1000 if sharded else ranks.rank_seed(1000, rank))
    jobs = 1 if sharded else world  # graphs finished per step by the whole job
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    tot_ms = timed(resident, args.steps, args.warmup)
    launches = (_lib.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if sampler else None
    e2e_ms = timed(e2e, args.steps, args.warmup)
    pipe_ms = None
    if not sharded:
        pipe_ms, _, _ = e2e_pipelined(args.steps, args.warmup)
    if sharded:
        from graphlearninglayer_b200 import sharded as sharded_mod

        info = sharded_mod.last_info()
    else:
        info = pkg.last_info()

    # forward / backward split, its own pass: two extra events per step
    split["ev"] = []
    for _ in range(args.steps):
        flush.zero_()
        resident()
    torch.cuda.synchronize()
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in split["ev"]) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in split["ev"]) / args.steps
    split["ev"] = None

    # per-kernel pass (instrumented; not the pass `value` comes from)
    _lib.lib.gll_profile_enable(1)
    _lib.profile_collect()
    for _ in range(args.steps):
        flush.zero_()
        resident()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.lib.gll_profile_enable(0)

    extra = {}
    if rank == 0 and world == 1 and args.workload not in ("c4",) + SHARDED and not args.no_large_graph:
        # the CG roofline study (BASELINE.json configs[3]); reported beside the headline, not instead of it
        r4, _, shp4, _, _, _, _ = step_fn("c4", 2000)
        ms4 = timed(r4, 5, 3)
        info4 = pkg.last_info()
        _lib.lib.gll_profile_enable(1)
        _lib.profile_collect()
        for _ in range(5):
            flush.zero_()
            r4()
        torch.cuda.synchronize()
        prof4 = _lib.profile_collect()
        _lib.lib.gll_profile_enable(0)
        extra["large_graph_c4"] = large_graph_report(prof4, shp4, info4, ms4 / 5)

    if rank == 0:
        peaks = load_peaks()
        per_step = tot_ms / args.steps
        kern = {}
        step_kernel_ms = sum(v[0] for v in prof.values()) / args.steps
        for nm, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            kern[nm] = {"ms_per_launch": ms / cnt, "launches_per_step": cnt / args.steps,
                        "share_of_kernel_time": (ms / args.steps) / step_kernel_ms}
        top = max(prof.items(), key=lambda kv: kv[1][0])[0]
        if sharded:
            top = "knn_gram_topk_tcgen05" if "knn_gram_topk_tcgen05" in prof else top
            shp = dict(shp, n_rows_this_rank=-(-shp["n"] // world))
        bound, work, unit = algorithmic_work(top, shp, info)
        if sharded and bound == "tensor":
            work = work / world  # each rank searches n/world rows against all n columns
        t_s = prof[top][0] / prof[top][1] * 1e-3
        peak_source = peaks["source"]
        if bound == "tensor":
            ach, peak, u = work / t_s / 1e12, peaks["tflops"], "TFLOP/s"
            if t_s > 0.05 and "tflops_sustained" in peaks:  # a launch of tens of ms runs under the power cap: sustained cuBLAS figure
                peak, peak_source = peaks["tflops_sustained"], "measured (MEASURED_PEAKS.json, sustained: the launch lasts > 50 ms)"
        else:
            ach, peak, u = (work / t_s / 1e9 if work else None), peaks["hbm_gbs"], "GB/s"
        traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same workload)
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(top)
        roofline = {"kernel": top, "bound": bound, "achieved": ach, "peak": peak, "unit": u,
                    "frac": (ach / peak if ach else None), "traffic": traffic, "peak_source": peak_source,
                    "algorithmic_work_per_launch": work, "work_unit": unit, "launch_ms": t_s * 1e3}
        if bound == "tensor" and top == "knn_gram_topk_tcgen05":
            # the Gram entry is accumulated from split operands: (hi + lo).hi in fp16 = 2 MMA passes (default), or
            # hi.hi + lo.hi + hi.lo in bf16 = 3 passes (GLL_B200_KNN_SPLIT=bf16x3); `achieved` counts 2 n^2 d once
            passes = 3 if os.environ.get("GLL_B200_KNN_SPLIT") == "bf16x3" else 2
            roofline.update({"mma_passes": passes, "issued": ach * passes, "issued_frac": ach * passes / peak})
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not sharded:
            times = cpu_calls(args.workload, 1000, 4, 25.0)
            cpu = {"value": len(times) / float(np.sum(times)), "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": f"{len(times)} full fwd+bwd calls of the same workload through oracle/gll_oracle.py (fp64)"}
        line = {
            "metric": METRIC, "value": jobs * 1e3 / per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic gaussian clusters (graphlearninglayer_b200.synth), L2-normalised, one graph per rank",
            "config": {"workload": args.workload + ": " + WORKLOAD_DESC[args.workload], "l2": "flushed between timed steps "
                       "(256 MiB memset outside the event pairs)", "parallelism": (f"one graph over {world} rank(s): rows (kNN, backward) + "
                                       + ("class columns (CG, no per-iteration collective)" if args.cg_partition == "columns" else
                                          "rows (CG: NCCL all-gather of the iterate + all-reduce of the dot products per iteration)"
                                          if args.cg_partition == "rows" else
                                          "rows (CG: iterate and dot products exchanged INSIDE the kernels over NVLink peer memory)")
                                       if sharded
                                       else f"independent graphs x{world}"),
                       "cg_tol": 1e-7,
                       "loss": ("custom_ce_loss as one kernel (graphlearninglayer_b200.losses; formula of losses.py:128-136)"
                                if args.loss == "fused" else "custom_ce_loss with the reference's PyTorch ops (losses.py:128-136)"),
                       "graph": {"nnz": info["nnz"], "nnz_uu": info["nnz_uu"],
                                                 "cg_iters_fwd": info["cg_iters_fwd"], "cg_iters_bwd": info["cg_iters_bwd"],
                                                 "knn_fallback_rows": info["knn_fallback_rows"], "status": info["status"],
                                                 **({"cg_solve_ms_fwd_bwd": info.get("cg_solve_ms")} if sharded else {})}},
            # e2e: host buffers in, host buffers out, every copy inside the timed region.  Headline = the double-buffered
            # pipeline (three streams; independent calls overlap their PCIe copies with the neighbours' kernels);
            # "serial" = one call at a time, copies and kernels back to back on one stream.
            "e2e": ({"value": jobs * 1e3 / (pipe_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": pipe_ms / args.steps,
                     "mode": "HostPipeline depth 3: H2D(i+1) | kernels(i) | D2H(i-1) on three streams; L2 not flushed "
                             "(inputs arrive by DMA every step, the in-flight steps touch ~290 MB > L2); bound by the "
                             "43.5 MB per step crossing PCIe in both directions at once (~54 GB/s in total)",
                     "serial": {"value": jobs * 1e3 / (e2e_ms / args.steps), "ms_per_step": e2e_ms / args.steps,
                                "mode": "one call at a time on one stream, L2 flushed between steps"}}
                    if pipe_ms is not None else
                    {"value": jobs * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps}),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "fwd_bwd_split_ms": {"forward_and_loss": fwd_ms, "backward": bwd_ms},
            "kernels": kern,
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def large_graph_report(prof, shp, info, ms_per_call):
    peaks = load_peaks()
    out = {"workload": "c4: " + WORKLOAD_DESC["c4"], "calls_per_sec": 1e3 / ms_per_call, "ms_per_call": ms_per_call,
           "graph": {k: info[k] for k in ("nnz", "nnz_uu", "cg_iters_fwd", "cg_iters_bwd", "cg_resid_fwd", "cg_resid_bwd",
                                          "knn_fallback_rows", "status")}}
    if "cg_persistent" in prof:
        ms, cnt = prof["cg_persistent"]
        iters = info["cg_iters_fwd"] + info["cg_iters_bwd"]
        solves_ms = ms / (cnt / 2)  # one fwd + one adjoint solve per call
        m, l, Euu = shp["m"], shp["l"], info["nnz_uu"]
        b_iter = 8.0 * Euu + 4.0 * (m + 1) + 44.0 * m * l
        gbs = b_iter * iters / (solves_ms * 1e-3) / 1e9
        out["cg"] = {"iters_per_sec": iters / (solves_ms * 1e-3), "us_per_iter": 1e3 * solves_ms / iters,
                     "bytes_per_iter": b_iter, "achieved_gbs": gbs, "peak_gbs": peaks["hbm_gbs"],
                     "frac": gbs / peaks["hbm_gbs"], "note": "working set is L2-resident at this size (SURVEY 8d)"}
    tot = sum(v[0] for v in prof.values())
    out["kernels"] = {nm: {"ms_per_launch": ms / cnt, "share_of_kernel_time": ms / tot}
                      for nm, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--loss", choices=["fused", "torch"], default="fused",
                    help="custom_ce_loss of the step: graphlearninglayer_b200.losses (one kernel) or the reference's PyTorch ops")
    ap.add_argument("--cg-partition", choices=["columns", "rows", "rows-p2p"], default="columns",
                    help="sharded workloads: split the CG solves by class columns (no per-iteration collective) or by rows "
                         "(all-gather + all-reduce per iteration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    os.environ["GLL_B200_SHARD_CG"] = args.cg_partition
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
