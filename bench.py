#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: RAFT correlation build + 12 lookups, frame-pairs/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: CorrBlock build (volume +
4-level pyramid) followed by `--iters` (12) lookups with fresh coordinates, for B=8 frame pairs of
436x1024 (padded 440x1024 -> 55x128 feature maps, D=256, r=4: BASELINE.json configs[1] shape) per GPU.
Work is sharded by frame pair across ranks with no collective on the data path (weak scaling:
every rank processes its own B pairs).  Prints ONE JSON line (rank 0).

`value`      : pairs/s with inputs already resident in HBM (CUDA events on the launch stream).
`e2e`        : the same step through the public API from pinned HOST buffers: H2D of both feature
               maps and all coordinate sets + D2H of the last lookup's features, inside the timing.
`roofline`   : the dominant kernel (largest share of the step), algorithmic bytes / measured time
               against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle's torch-CPU port of the reference CorrBlock (same ATen ops the reference
               calls) on this box's host cores, bounded sample (B=1).
`--impl reference` times that CPU port alone (the reference's Python cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RAFT corr build+12-iter lookup frame-pairs/s @436x1024"
UNIT = "frame-pairs/s"
SHAPE = dict(B=8, D=256, H=55, W=128, L=4, r=4)  # per GPU


def algorithmic_bytes(D, H, W, L, r):
    """SURVEY.md section 8(d) / BASELINE.md section 2, per frame pair."""
    N = H * W
    K = 2 * r + 1
    sizes, h, w = [], H, W
    for _ in range(L):
        sizes.append(h * w)
        h, w = h // 2, w // 2
    build = 2 * N * D * 4 + 4 * N * sum(sizes)
    lookup = N * (L * (K + 1) ** 2 * 4 + L * K * K * 4 + 8)
    flops = 2 * N * N * D
    return build, lookup, flops


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 0.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Index of the next sample: brackets a window of the run for :meth:`stats`."""
        return len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.stats(0, len(self.lines))

    def stats(self, lo, hi):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        power.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w": power[len(power) // 2] if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arms
def staged_reference_corr():
    """The reference's own, unmodified ``raft/core/corr.py`` (+ ``core/utils/utils.py``) from the git-ignored
    ``baseline/_ref/stock`` tree that ``build()`` stages and the snapshot carries to the GPU box
    (scripts/vendor_reference.py); ``None`` when it is not there.  Baseline / checker only -- never on the product path."""
    base = os.path.join(ROOT, "baseline", "_ref", "stock")
    if not os.path.exists(os.path.join(base, "raft", "core", "corr.py")):
        return None
    import importlib

    for p in (os.path.join(base, "raft"), base):  # `core.*`, and the reference's own compiled alt_cuda_corr*.so
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("core.corr")
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(base)), mod.__file__
    return mod


def cpu_port_run(steps, warmup, iters, shape, seed=1234):
    """Times the reference CorrBlock on the host cores, one step = one frame pair (B=1): the reference's own file
    from baseline/_ref (``kind: "reference"``) when staged, else the oracle's torch-CPU port of it (``"port"``)."""
    import torch

    ref = staged_reference_corr()
    if ref is not None:
        CpuCorrBlock, kind = ref.CorrBlock, "reference"
    else:
        from oracle.torch_cpu_port import CpuCorrBlock
        kind = "port"

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, H, W, L, r = shape["D"], shape["H"], shape["W"], shape["L"], shape["r"]
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn((1, D, H, W), generator=g)
    f2 = torch.randn((1, D, H, W), generator=g)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None]
    coords = [grid + 4.0 * torch.randn((1, 2, H, W), generator=g)]
    for _ in range(iters - 1):
        coords.append(coords[-1] + 0.5 * torch.randn((1, 2, H, W), generator=g))
    times = []
    with torch.no_grad():
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            blk = CpuCorrBlock(f1, f2, L, r)
            for c in coords:
                out = blk(c)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    _ = float(out.sum())
    total = sum(times)
    what = ("the reference's unmodified raft/core/corr.py CorrBlock (baseline/_ref/stock)" if kind == "reference"
            else "oracle/torch_cpu_port.py (same three ATen ops)")
    return {"pairs_per_s": len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"B=1 frame pair x {len(times)} steps (build + {iters} lookups, D={D}, {H}x{W}, r={r}), {what}, torch {torch.__version__} CPU ops"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    r = cpu_port_run(args.steps, max(args.warmup, 1), args.iters, SHAPE)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"RAFT-basic CorrBlock build + {args.iters} lookups, 436x1024 (fmap 55x128, D=256, r=4, L=4), "
                               "1 frame pair per step on host cores", "iters": args.iters},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ---------------------------------------------------------------------------------------------- GPU arm
def pin_to_gpu_cpus(gpu_index):
    """Restrict this rank to the CPUs nvidia-smi lists as local to its GPU (same NUMA node / PCIe root), BEFORE the
    pinned host buffers are allocated (first touch puts them on that node): with 8 ranks copying at once the default
    placement sends most H2D traffic across the socket link.  Best effort: returns a description or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        head = None
        for ln in out.splitlines():
            cols = [c.strip() for c in ln.split("\t")]
            if head is None and any("CPU Affinity" in c for c in cols):
                head = [c for c in cols]
                continue
            if head is not None and cols and cols[0] == f"GPU{gpu_index}":
                # the header has one leading empty cell per row label
                idx = next(i for i, c in enumerate(head) if "CPU Affinity" in c)
                spec = cols[idx] if idx < len(cols) else ""
                cpus = set()
                for part in spec.split(","):
                    part = part.strip()
                    if not part:
                        continue
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    return f"{spec} ({len(cpus)} CPUs)"
                return None
    except Exception:
        return None
    return None


def alt_1080p_side(rc, torch, dev, g, e0, e1, hbm, with_reference):
    """north_star config 5a: AlternateCorrBlock (on-the-fly correlation) lookups at 1080x1920 -> 135x240, D=256, r=4,
    L=4, B=4.  Algorithmic work per pair per lookup: bytes 4 D (N + sum H_l W_l) + 8 N + 4 N L K^2,
    flops N L (K+1)^2 2 D."""
    B, D, H, W, L, r, n_coords = 4, 256, 135, 240, 4, 4, 4
    K = 2 * r + 1
    N = H * W
    sizes, h, w = [], H, W
    for _ in range(L):
        sizes.append(h * w)
        h, w = h // 2, w // 2
    alg_bytes = B * (4 * D * (N + sum(sizes)) + 8 * N + 4 * N * L * K * K)
    flops = B * N * L * (K + 1) ** 2 * 2 * D
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = 4.0 * torch.randn((B, 2, H, W), generator=g, device=dev)
    flow = torch.nn.functional.avg_pool2d(flow, 5, stride=1, padding=2, count_include_pad=False) * 5.0
    coords = [(grid + flow).contiguous()]
    for _ in range(n_coords - 1):
        coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())

    def timed(fn, reps):
        for c in coords:
            fn(c)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            e0.record()
            for c in coords:
                fn(c)
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1) / len(coords))
        return sorted(ts)[len(ts) // 2]

    blk = rc.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
    ms = timed(blk, 5)
    res = {"ms_per_lookup": ms, "algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / hbm, "useful_tflops": flops / (ms * 1e-3) / 1e12,
           "pair_lookups_per_s": B / (ms * 1e-3),
           "what": f"AlternateCorrBlock.__call__ at 1080p (fmap {H}x{W}, D={D}, r={r}, L={L}, B={B}), smooth sigma-4px flow, median of 5 x {n_coords} lookups; "
                   "north_star config 5a, not part of the headline step"}
    if with_reference:
        ref_mod = staged_reference_corr()
        try:
            if ref_mod is None:
                raise RuntimeError("baseline/_ref/stock is not staged")
            import alt_cuda_corr  # noqa: F401  the reference's own extension, built from its sources for sm_100 by build()
            rblk = ref_mod.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)
            res["reference_alt_cuda_corr_ms_per_lookup"] = timed(rblk, 2)
            res["reference_max_abs_delta_over_max_abs"] = float((rblk(coords[0]) - blk(coords[0])).abs().max() / rblk(coords[0]).abs().max())
            res["reference_what"] = ("the reference's unmodified AlternateCorrBlock (corr.py:119-167) over its own alt_cuda_corr kernel "
                                     "(correlation_kernel.cu compiled for sm_100), same GPU and inputs")
        except Exception as exc:  # the extension is a baseline only: report why it is absent, never fail the bench
            res["reference_alt_cuda_corr_ms_per_lookup"] = None
            res["reference_unavailable"] = f"{type(exc).__name__}: {exc}"[:200]
    return res


def lookup_ablation_side(torch, dev, blk, coords, e0, e1, l_ms):
    """What the memory system charges for the lookup's box list, measured in this run: the production
    kernel with its evaluation compiled out (RAFTCORR_LOOKUP_EXP=1: same TMA boxes, same waits, nothing evaluated or
    written) and with its boxes compiled out (=2: evaluation + write-out from stale shared memory).  The switch is read
    per launch; the outputs of these launches are NOT lookups and are discarded."""
    res = {}
    try:
        for key, exp in (("gather_only_ms", "1"), ("evaluate_and_store_only_ms", "2")):
            os.environ["RAFTCORR_LOOKUP_EXP"] = exp
            for c in coords:
                blk(c)
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(5):
                e0.record()
                for c in coords:
                    blk(c)
                e1.record()
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1) / len(coords))
            res[key] = sorted(ts)[2]
    finally:
        os.environ.pop("RAFTCORR_LOOKUP_EXP", None)
    res["frac_of_gather_peak"] = res["gather_only_ms"] / l_ms
    res["what"] = ("timing-only ablations of the same kernel on the same coordinates: frac_of_gather_peak = time of the gather "
                   "alone / time of the lookup (1.0 = the lookup costs what the memory system charges for its boxes)")
    return res


def train_step_side(rc, torch, dev, g):
    """BASELINE.json configs[2] on one GPU: B=10 pairs at 368x496 (fmap 46x62, D=256, r=4, L=4), forward + backward
    through 12 lookups and the volume (`out.backward(randn)` per iteration, accumulated: SURVEY 8d), replayed as ONE CUDA
    graph -- only this library's kernels in the step.  Beside it the same step with the per-iteration lookup adjoint
    (red.global.add + fold passes, round 1's path)."""
    B, D, H, W, L, r, iters = 10, 256, 46, 62, 4, 4, 12
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
    coords = [(grid + flow).contiguous()]
    for _ in range(iters - 1):
        coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
    ws = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), generator=g, device=dev) for _ in range(iters)]
    a, b = f1.clone().requires_grad_(), f2.clone().requires_grad_()

    def step():
        a.grad = None
        b.grad = None
        blk = rc.CorrBlock(a, b, L, r)
        torch.autograd.backward([blk(c) for c in coords], ws)

    def graph_ms():
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):  # NCCL's watchdog thread may touch the device
            step()
        gr.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            gr.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / 20

    res = {"what": f"config 3: B={B}, fmap {H}x{W}, D={D}, r={r}, L={L}, {iters} iterations, forward + backward as one CUDA graph "
                   "(autograd.backward(outs, randn weights)); not part of the headline step"}
    old = os.environ.get("RAFTCORR_LOOKUP_BWD")
    try:
        os.environ["RAFTCORR_LOOKUP_BWD"] = "batched"
        res["ms_per_step"] = graph_ms()
        res["pairs_per_s"] = B / (res["ms_per_step"] * 1e-3)
        os.environ["RAFTCORR_LOOKUP_BWD"] = "immediate"
        res["ms_per_step_per_iteration_adjoint"] = graph_ms()
    finally:
        if old is None:
            os.environ.pop("RAFTCORR_LOOKUP_BWD", None)
        else:
            os.environ["RAFTCORR_LOOKUP_BWD"] = old
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import optical_flow_dnn_b200 as rc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = None if args.no_affinity else pin_to_gpu_cpus(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.l2_fetch:
        import ctypes

        from optical_flow_dnn_b200 import _lib
        prev = ctypes.c_int(0)
        _lib.check(_lib.lib().raftcorr_set_l2_fetch_granularity(args.l2_fetch, local_rank, ctypes.byref(prev)), "l2 fetch")
        if rank == 0:
            print(f"[bench] L2 fetch granularity {prev.value} -> {args.l2_fetch} bytes", file=sys.stderr)
    B, D, H, W, L, r = (SHAPE[k] for k in "BDHWLr")
    B = args.batch or B
    iters = args.iters
    K = 2 * r + 1
    mode = args.mode
    # the job is world x B frame pairs, sharded by pair (optical_flow_dnn_b200/sharding.py: torch.chunk ranges, no
    # collective); every pair is generated from its GLOBAL index, so a rank's data does not depend on the world size
    from optical_flow_dnn_b200.sharding import shard_range
    mine = shard_range(world * B, rank, world)
    assert len(mine) == B
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()
    f1 = torch.empty((B, D, H, W), device=dev)
    f2 = torch.empty((B, D, H, W), device=dev)
    coords = [torch.empty((B, 2, H, W), device=dev) for _ in range(iters)]
    for i, pair in enumerate(mine):
        g = torch.Generator(device=dev).manual_seed(1234 + pair)
        f1[i] = torch.randn((D, H, W), generator=g, device=dev)
        f2[i] = torch.randn((D, H, W), generator=g, device=dev)
        # "realistic" flow: sigma 4 px at 1/8 res smoothed by a 5x5 box filter, fresh coords per iteration
        flow = 4.0 * torch.randn((1, 2, H, W), generator=g, device=dev)
        flow = torch.nn.functional.avg_pool2d(flow, 5, stride=1, padding=2, count_include_pad=False)[0] * 5.0
        coords[0][i] = grid + flow
        for k in range(1, iters):
            coords[k][i] = coords[k - 1][i] + 0.5 * torch.randn((2, H, W), generator=g, device=dev)
    g = torch.Generator(device=dev).manual_seed(99 + rank)  # side measurements (weights etc.)

    def step():
        blk = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
        out = None
        for c in coords:
            out = blk(c)
        return blk, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        # start the clock sampler BEFORE the warm-up: nvidia-smi's start-up (NVML init) takes ~100 ms and
        # contends for the driver; only its steady 100 ms polling should overlap the timed region
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(max(args.warmup, 3)):
            blk, out = step()
        if rank == 0:
            t_wait = time.time()
            while sampler.proc and not sampler.lines and time.time() - t_wait < 3.0:
                time.sleep(0.01)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            blk, out = step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        # per-kernel shares, measured live with events on the launch stream (same inputs), right behind the timed region --
        # i.e. under the clocks the headline ran at, not behind the 2 s power-capped sustained loop (where round 1 took them)
        kb0, kb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kl0, kl1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        build_ms, look_ms = [], []
        for _ in range(max(3, min(args.steps, 10))):
            kb0.record()
            blk = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
            kb1.record()
            kl0.record()
            for c in coords:
                out = blk(c)
            kl1.record()
            torch.cuda.synchronize(dev)
            build_ms.append(kb0.elapsed_time(kb1))
            look_ms.append(kl0.elapsed_time(kl1) / iters)
            del blk
        l_ms_now = sorted(look_ms)[len(look_ms) // 2]
        ablation = None
        if rank == 0 and mode != "fp32" and not args.no_fused and os.environ.get("RAFTCORR_PYRAMID_LAYOUT", "quads") == "quads":
            try:  # a side line must never take the headline down with it
                blk = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
                ablation = lookup_ablation_side(torch, dev, blk, coords, kl0, kl1, l_ms_now)
                del blk
            except Exception as exc:
                ablation = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        # ---- sustained: the same step back to back for >= 2 s (the build runs into the board's power cap; the
        # headline above is a short burst) with clocks and power sampled over exactly this window
        sustained = None
        if not args.no_sustained:
            # the same number of steps on every rank (sized from the slowest rank's headline time)
            n_t = torch.tensor([math.ceil(args.sustained_s * 1e3 / (ms_total / args.steps))], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(n_t, op=dist.ReduceOp.MIN)
            n_sus = int(n_t[0])
            m0 = sampler.mark() if rank == 0 else 0
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i_ in range(n_sus):
                blk, out = step()
                if i_ % 64 == 63:
                    torch.cuda.synchronize(dev)  # keep the caching allocator's pending frees bounded
            s1.record()
            barrier()
            sus_ms = s0.elapsed_time(s1)
            sustained = {"steps": n_sus, "ms_per_step": sus_ms / n_sus, "seconds": sus_ms * 1e-3}
            if rank == 0:
                st = sampler.stats(m0, sampler.mark())
                sustained.update({"sm_mhz": st["sm_mhz"], "power_w": st["power_w"], "reasons": st["reasons"], "clock_samples": st["samples"]})
        # ---- the same step with TF32 operands (SURVEY K1's MMA kind), beside the f16 default
        tf32_step = None
        if rank == 0 and mode == "f16" and not args.no_fused:
            def step_tf32():
                b_ = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode="tf32")
                for c in coords:
                    o_ = b_(c)
                return o_
            for _ in range(3):
                step_tf32()
            torch.cuda.synchronize(dev)
            kb0.record()
            for _ in range(10):
                step_tf32()
            kb1.record()
            torch.cuda.synchronize(dev)
            tf32_step = {"ms_per_step": kb0.elapsed_time(kb1) / 10, "pairs_per_s": B * 10 / (kb0.elapsed_time(kb1) * 1e-3),
                         "what": "build_mode='tf32' (kind::tf32 MMAs, TF32-rounded operands), same step, 10 steps after 3 warm-up"}
        # ---- precision, measured in this run: 4096 random level-0 volume entries against fp64 dot products of the same
        # features, and one full lookup against the exact-fp32 build mode's lookup
        precision = None
        if rank == 0:
            blk = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
            lvl0 = blk.corr_pyramid[0].view(B, H * W, H * W)
            gi = torch.Generator(device=dev).manual_seed(7)
            bi = torch.randint(0, B, (4096,), generator=gi, device=dev)
            mi = torch.randint(0, H * W, (4096,), generator=gi, device=dev)
            ni = torch.randint(0, H * W, (4096,), generator=gi, device=dev)
            got = lvl0[bi, mi, ni].double()
            a_ = f1.view(B, D, H * W)[bi, :, mi].double()
            b_ = f2.view(B, D, H * W)[bi, :, ni].double()
            ref = (a_ * b_).sum(1) / math.sqrt(D)
            vmax = float(lvl0.abs().max())
            out_m = blk(coords[0])
            del lvl0, blk
            out_x = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode="fp32")(coords[0])
            precision = {"measured": {"volume_max_abs_err_over_max_abs": float((got - ref).abs().max()) / vmax,
                                      "volume_entries_checked": 4096, "volume_reference": "fp64 dot products of the same features / sqrt(D)",
                                      "lookup_max_abs_err_over_max_abs_vs_fp32_mode": float((out_m - out_x).abs().max() / out_x.abs().max())},
                         "tolerance": 1e-3, "build_mode": mode}
            del out_m, out_x
        # ---- like-for-like GPU baseline: the reference's own CorrBlock (stock PyTorch CUDA ops: cuBLAS SGEMM, avg_pool2d,
        # grid_sample) on the same GPU and inputs -- baseline/_ref/stock/raft/core/corr.py, unmodified
        gpu_baseline = None
        if rank == 0 and not args.no_gpu_baseline:
            ref_mod = staged_reference_corr()
            if ref_mod is None:
                gpu_baseline = {"unavailable": "baseline/_ref/stock is not staged (build() stages it where /root/reference exists)"}
            else:
                tf32_before = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False  # the reference's default: fp32 SGEMM

                def step_ref():
                    b_ = ref_mod.CorrBlock(f1, f2, num_levels=L, radius=r)
                    for c in coords:
                        o_ = b_(c)
                    return o_
                o_ref = step_ref()
                o_our = step()[1]
                agree = float((o_our - o_ref).abs().max() / o_ref.abs().max())
                del o_ref, o_our
                step_ref()
                torch.cuda.synchronize(dev)
                kb0.record()
                for _ in range(5):
                    step_ref()
                kb1.record()
                torch.cuda.synchronize(dev)
                torch.backends.cuda.matmul.allow_tf32 = tf32_before
                ms_ref = kb0.elapsed_time(kb1) / 5
                gpu_baseline = {"value": B * 1e3 / ms_ref, "unit": UNIT, "ms_per_step": ms_ref, "kind": "reference",
                                "what": "the reference's unmodified raft/core/corr.py CorrBlock (baseline/_ref/stock) with stock PyTorch "
                                        f"{torch.__version__} CUDA kernels, fp32 (allow_tf32 off), same GPU, same inputs, same step (B={B}), 5 steps after 2 warm-up",
                                "last_lookup_max_abs_delta_over_max_abs_vs_this_library": agree}
        # ---- the on-the-fly class (north_star config 5a): AlternateCorrBlock lookups at 1080p, B=4, beside the headline
        alt = None
        if rank == 0 and not args.no_fused:
            alt = alt_1080p_side(rc, torch, dev, g, kl0, kl1, peaks()[0], not args.no_gpu_baseline)
        # ---- north_star config 3 on this GPU: the training step (forward + backward) as one CUDA graph
        train = None
        if rank == 0 and not args.no_fused:
            torch.cuda.empty_cache()
            try:
                with torch.enable_grad():
                    train = train_step_side(rc, torch, dev, g)
            except Exception as exc:
                train = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        # beside the headline: the same build with the three-term fp16 split (fp32-class accuracy) and, for scale, what
        # the exact-fp32 FFMA build costs
        accurate = None
        if rank == 0 and mode != "fp32" and not args.no_fused:
            accurate = {}
            for m_ in ("f16x3", "fp32"):
                rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=m_)
                torch.cuda.synchronize(dev)
                kb0.record()
                for _ in range(3):
                    rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=m_)
                kb1.record()
                torch.cuda.synchronize(dev)
                accurate[m_ + "_build_ms"] = kb0.elapsed_time(kb1) / 3
            accurate["what"] = ("build_mode='f16x3': [hi|lo|hi] x [hi|hi|lo] fp16 split through the same tensor-core kernel, "
                                "max|dV|/max|V| 2.5e-6; 'fp32': exact FFMA build on the CUDA cores (8.5e-7)")
        # SURVEY.md 8(f) rank 1, reported beside the headline (not part of it): the lookup fused with the motion
        # encoder's first 1x1 convolution + ReLU (update.py:158,170: 324 -> 256) against the unfused pair
        # (our lookup followed by cuDNN's convolution), per refinement iteration, same coordinates
        fused = None
        if rank == 0 and mode != "fp32" and not args.no_fused:
            cw = torch.randn((256, L * K * K, 1, 1), generator=g, device=dev) / math.sqrt(L * K * K)
            cb = torch.randn((256,), generator=g, device=dev)
            blk = rc.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)

            def per_iter(fn):
                for c in coords:
                    fn(c)
                torch.cuda.synchronize(dev)
                ts = []
                for _ in range(5):
                    kl0.record()
                    for c in coords:
                        fn(c)
                    kl1.record()
                    torch.cuda.synchronize(dev)
                    ts.append(kl0.elapsed_time(kl1) / iters)
                return sorted(ts)[len(ts) // 2]

            f_ms = per_iter(lambda c: blk.lookup_convc1(c, cw, cb))
            u_ms = per_iter(lambda c: torch.relu_(torch.nn.functional.conv2d(blk(c), cw, cb)))
            fused = {"ms_per_launch": f_ms, "unfused_lookup_plus_cudnn_conv1x1_relu_ms": u_ms, "cout": 256,
                     "what": "CorrBlock.lookup_convc1 (lookup -> TF32 tcgen05 product with convc1's weight -> bias, ReLU); "
                             "SURVEY 8(f) rank 1, not part of the headline step"}
            del blk
        # SURVEY.md 8(f) rank 2, also beside the headline: the encoders' 1x1 conv2 head (128 -> 256) fused in front of
        # the build (CorrBlock.from_encoder_head) against cuDNN's conv2 + split + CorrBlock(fmap1, fmap2)
        head = None
        if rank == 0 and mode != "fp32" and not args.no_fused:
            hx = [torch.randn((2 * B, 128, H, W), generator=g, device=dev) for _ in range(2)]
            hw = torch.randn((D, 128, 1, 1), generator=g, device=dev) / math.sqrt(128.0)
            hb = 0.1 * torch.randn((D,), generator=g, device=dev)

            def per_call(fn):
                for i in range(3):
                    fn(hx[i & 1])
                torch.cuda.synchronize(dev)
                ts = []
                for _ in range(5):
                    kl0.record()
                    for i in range(4):
                        fn(hx[i & 1])
                    kl1.record()
                    torch.cuda.synchronize(dev)
                    ts.append(kl0.elapsed_time(kl1) / 4)
                return sorted(ts)[len(ts) // 2]

            def unfused_head(x_):
                f_ = torch.nn.functional.conv2d(x_, hw, hb)
                return rc.CorrBlock(f_[:B], f_[B:], num_levels=L, radius=r, build_mode=mode)

            head = {"ms_per_call": per_call(lambda x_: rc.CorrBlock.from_encoder_head(x_, hw, hb, num_levels=L, radius=r)),
                    "cudnn_conv2_plus_build_ms": per_call(unfused_head), "cin": 128,
                    "what": "CorrBlock.from_encoder_head (tcgen05 TF32 1x1 conv writing the build's operand buffers, then the "
                            "build); SURVEY 8(f) rank 2, not part of the headline step"}
        clocks = sampler.stop() if rank == 0 else None

        # ---- e2e: public API from pinned host buffers, copies inside the timed region.
        # Every step copies its own inputs host->device and its result device->host; consecutive steps are
        # software-pipelined over three streams (H2D | compute | D2H, two buffer sets) the way a serving loop
        # would, so PCIe transfers of step i+1 / i-1 overlap the kernels of step i.
        hf1, hf2 = f1.cpu().pin_memory(), f2.cpu().pin_memory()
        hco = torch.stack([c.cpu() for c in coords]).pin_memory()
        houts = [torch.empty((B, L * K * K, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
        h2d = (hf1.numel() + hf2.numel() + hco.numel()) * 4
        d2h = houts[0].numel() * 4
        s_in, s_cmp, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        dbuf = [(torch.empty_like(f1), torch.empty_like(f2), torch.empty((iters, B, 2, H, W), device=dev)) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n):
            for i in range(n):
                k = i & 1
                d1, d2, dc = dbuf[k]
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_cmp[k])   # the kernels of step i-2 are done with this buffer set
                    d1.copy_(hf1, non_blocking=True)
                    d2.copy_(hf2, non_blocking=True)
                    dc.copy_(hco, non_blocking=True)
                    ev_in[k].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[k])
                    b_ = rc.CorrBlock(d1, d2, num_levels=L, radius=r, build_mode=mode)
                    o = None
                    for j in range(iters):
                        o = b_(dc[j])
                    ev_cmp[k].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[k])
                    s_out.wait_event(ev_out[k])  # host buffer k was last filled two steps ago
                    houts[k].copy_(o, non_blocking=True)
                    o.record_stream(s_out)
                    ev_out[k].record(s_out)

        e2e_run(4)
        # what the link gives a lone pinned copy on this box, to read the e2e number against (rank 0's link)
        for st in (s_in, s_cmp, s_out):
            st.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s_in):
            dbuf[0][0].copy_(hf1, non_blocking=True)
            p0.record(s_in)
            for _ in range(4):
                dbuf[0][0].copy_(hf1, non_blocking=True)
            p1.record(s_in)
        s_in.synchronize()
        h2d_alone_gbs = 4 * hf1.numel() * 4 / (p0.elapsed_time(p1) * 1e-3) / 1e9
        # The copies cross a PCIe link / host memory system shared with whatever else runs on the box: the same code
        # measures 2.5 ms and 4.8 ms per step minutes apart.  Three trials of K steps each, the MEDIAN is reported
        # (all three are in the JSON line).
        e2e_trials = []
        for _ in range(3):
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(s_in)
            e2e_run(args.steps)
            for st in (s_in, s_cmp):
                s_out.wait_stream(st)
            x1.record(s_out)
            barrier()
            e2e_trials.append(x0.elapsed_time(x1))
        e2e_ms = sorted(e2e_trials)[1]
        # ---- the platform's ceiling for this step's host traffic: the SAME copies (pinned H2D of both feature maps and
        # the coordinates, D2H of the result) on the same streams with NO kernels, all ranks at once
        def copies_only(n):
            for i in range(n):
                k = i & 1
                d1, d2, dc = dbuf[k]
                with torch.cuda.stream(s_in):
                    d1.copy_(hf1, non_blocking=True)
                    d2.copy_(hf2, non_blocking=True)
                    dc.copy_(hco, non_blocking=True)
                with torch.cuda.stream(s_out):
                    houts[k].copy_(dout, non_blocking=True)
        dout = torch.empty((B, L * K * K, H, W), dtype=torch.float32, device=dev)
        copies_only(2)
        for st in (s_in, s_out):
            st.synchronize()
        copy_trials = []
        for _ in range(3):
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(s_in)
            copies_only(args.steps)
            s_out.wait_stream(s_in)
            x1.record(s_out)
            barrier()
            copy_trials.append(x0.elapsed_time(x1))
        copy_ms = sorted(copy_trials)[1]
        del dout

    t = torch.tensor([ms_total, e2e_ms, sus_ms if sustained else 0.0, copy_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, sus_ms_max, copy_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    line = None
    if rank == 0:
        hbm, bf16, peak_src = peaks()
        bb, lb, flops = algorithmic_bytes(D, H, W, L, r)
        b_ms = sorted(build_ms)[len(build_ms) // 2]
        l_ms = sorted(look_ms)[len(look_ms) // 2]
        build_share = b_ms / (b_ms + iters * l_ms)
        kern = {
            "build": {"ms_per_launch_group": b_ms, "achieved_gbs": B * bb / (b_ms * 1e-3) / 1e9,
                      "achieved_tflops": B * flops / (b_ms * 1e-3) / 1e12, "share_of_step": build_share},
            "lookup": {"ms_per_launch": l_ms, "achieved_gbs": B * lb / (l_ms * 1e-3) / 1e9,
                       "share_of_step": 1.0 - build_share},
        }
        # the step as a whole against the HBM roofline: algorithmic bytes of build + all lookups over the timed step
        step_gbs = B * (bb + iters * lb) / (ms_total / args.steps * 1e-3) / 1e9
        kern["step"] = {"algorithmic_bytes": B * (bb + iters * lb), "achieved_gbs": step_gbs, "frac_of_hbm_peak": step_gbs / hbm}
        if tf32_step:
            kern["step_tf32_mode"] = tf32_step
        if alt:
            kern["alt_1080p"] = alt
        if train:
            kern["train_step_config3"] = train
        if accurate:
            kern["accurate_build_modes"] = accurate
        if fused:
            kern["lookup_convc1"] = fused
        if head:
            kern["fnet_head_build"] = head
        lookup_name = ("lookup_fwd_kernel<4,4>" if os.environ.get("RAFTCORR_LOOKUP") == "cpasync" else
                       "lookup_fwd_tma_kernel<4,4>" if mode == "fp32" or os.environ.get("RAFTCORR_PYRAMID_LAYOUT", "quads") != "quads"
                       else "lookup_fwd_quad_kernel<4,4>")
        if build_share >= 0.5:
            dom, ach, alg = ("build_tc_kernel" if mode != "fp32" else "sgemm_kernel"), kern["build"]["achieved_gbs"], B * bb
        else:
            dom, ach, alg = lookup_name, kern["lookup"]["achieved_gbs"], B * lb
        traffic = None
        try:  # per-launch DRAM bytes of that kernel from the committed ncu --set full capture (same B, shape)
            if B == SHAPE["B"]:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        except Exception:
            traffic = None
        # tf32: 1 paired NCHW->NHWC transpose + build_tc; fp32: sgemm + (L-1) pooling passes; + one lookup per iteration
        launches_per_step = (2 if mode != "fp32" else 1 + (L - 1)) + iters
        cpu = cpu_port_run(3, 1, iters, SHAPE) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"tf32": "tf32", "f16": "f16xf16->f32", "fp32": "f32"}[mode],
            "data": "synthetic",
            "precision": precision,
            "config": {"workload": f"RAFT-basic CorrBlock build + {iters} lookups @436x1024 (fmap {H}x{W}, D={D}, r={r}, L={L}), "
                                   f"B={B} frame pairs per GPU (BASELINE.json configs[1] shape)",
                       "pairs_per_gpu": B, "iters": iters, "build_mode": mode, "parallelism": f"pair-sharded x{world}, no collective",
                       "l2": f"pyramid {B * 4 * H * W * (H * W) * 1.33 / 1e9:.2f} GB per GPU >> 126 MB L2 (inputs larger than L2)"},
            "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "h2d_gbs_in_step": h2d / (e2e_ms / args.steps * 1e-3) / 1e9, "h2d_gbs_lone_pinned_copy": h2d_alone_gbs,
                    "trials_ms_per_step": [t_ / args.steps for t_ in e2e_trials], "reported": "median of 3 trials of K steps (rank 0's trials listed)",
                    "copies_only": {"value": world * B * args.steps / (copy_ms * 1e-3), "unit": UNIT, "ms_per_step": copy_ms / args.steps,
                                    "aggregate_h2d_gbs": world * h2d / (copy_ms / args.steps * 1e-3) / 1e9,
                                    "aggregate_d2h_gbs": world * d2h / (copy_ms / args.steps * 1e-3) / 1e9,
                                    "what": "the platform ceiling for this step: the same pinned H2D + D2H copies on the same streams "
                                            "with no kernels, all ranks at once (median of 3, max over ranks)"},
                    "frac_of_copies_only": (copy_ms / e2e_ms),
                    "cpu_affinity": affinity,
                    "d2h_covers": f"last of {iters} lookup outputs ({d2h} of {iters * d2h} bytes produced per step)",
                    "result": "last lookup's [B,324,H,W] features copied to pinned host memory each step; "
                              "steps pipelined over H2D | compute | D2H streams (2 buffer sets)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         "traffic": traffic, "traffic_source": "static: profiles/ncu_traffic.json (ncu --set full capture of this "
                         "kernel at this shape, committed; not measured in this run)" if traffic is not None else None,
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src},
            "kernels": kern,
        }
        if ablation:
            line["roofline"]["lookup_ablation"] = ablation
        if sustained:
            sustained["value"] = world * B * sustained["steps"] / (sus_ms_max * 1e-3)
            sustained["unit"] = UNIT
            sustained["what"] = f">= {args.sustained_s:g} s of back-to-back steps after the timed region; device-timed, max over ranks; clocks/power: rank 0's GPU, median over the window"
            line["sustained"] = sustained
        if gpu_baseline:
            line["gpu_baseline"] = gpu_baseline
        if cpu:
            line["cpu_baseline"] = {"value": cpu["pairs_per_s"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                                    "sample": cpu["sample"]}
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


class _StdoutToStderr:
    """Everything a library prints on fd 1 while we run (e.g. NCCL's version banner) goes to stderr, so that
    stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=12, help="lookups per build (metric: 12)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("RAFTCORR_BUILD_MODE", "f16"), choices=["tf32", "f16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="frame pairs per GPU (default 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="skip the side measurements (lookup+convc1, encoder head, on-the-fly 1080p, tf32 step)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch reference CorrBlock on the GPU")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="do not pin the rank to its GPU's local CPUs")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="length of the sustained loop (seconds)")
    ap.add_argument("--l2-fetch", type=int, default=0, help="set cudaLimitMaxL2FetchGranularity (32/64/128) first")
    args = ap.parse_args()
    line = None
    with _StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
