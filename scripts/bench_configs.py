"""Secondary measurements: the other BASELINE.json configs (parity-test shapes) timed with CUDA events.
One JSON line per config -> profiles/. Not the headline bench (bench.py).

Single GPU: `python scripts/bench_configs.py [1 2 3 4 5a 5b]`.  N GPUs of one box (weak scaling: every rank runs the
config on its own seeded pairs, no collective on the path; barrier + max-over-ranks device time, whole-job pairs/s):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_configs.py 1 2 3 4 5a 5b
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import optical_flow_dnn_b200 as rc

WORLD = int(os.environ.get("WORLD_SIZE", "1")); RANK = int(os.environ.get("RANK", "0")); LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)
if WORLD > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
_print = print


def print(line):  # rank 0 reports; every line carries the GPU count
    if RANK == 0:
        d = json.loads(line)
        d["n_gpus"] = WORLD
        for k in ("pairs_per_s", "pair_iters_per_s"):
            if k in d:
                d[k] *= WORLD  # whole-job throughput (weak scaling: per-rank work is fixed)
        _print(json.dumps(d), flush=True)


def _max_over_ranks(ms):
    if WORLD == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

HBM = 6547.2
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def alg_bytes(D, H, W, L, r):
    N, K = H * W, 2 * r + 1
    sizes, h, w = [], H, W
    for _ in range(L):
        sizes.append(h * w); h, w = h // 2, w // 2
    build = 2 * N * D * 4 + 4 * N * sum(sizes)
    lookup = N * (L * (K + 1) ** 2 * 4 + L * K * K * 4 + 8)
    lookup_bwd = N * (L * K * K * 4 + 8 + L * (K + 1) ** 2 * 8)
    alt = 4 * D * (N + sum(sizes)) + 8 * N + 4 * N * L * K * K
    alt_flops = N * L * (K + 1) ** 2 * 2 * D
    return build, lookup, lookup_bwd, alt, alt_flops, sum(sizes)


def inputs(B, D, H, W, iters, seed=1234):
    g = torch.Generator(device=dev).manual_seed(seed + RANK)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
    coords = [(grid + flow).contiguous()]
    for _ in range(iters - 1):
        coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
    return f1, f2, coords


def timeit(fn, reps, warm=3):
    for _ in range(warm): fn()
    if WORLD > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    if WORLD > 1: dist.barrier()
    torch.cuda.synchronize()
    return _max_over_ranks(e0.elapsed_time(e1) / reps)


def fwd_config(name, B, D, H, W, L, r, iters, mode="tf32", reps=20):
    f1, f2, coords = inputs(B, D, H, W, iters)
    def step():
        blk = rc.CorrBlock(f1, f2, L, r, build_mode=mode)
        for c in coords: o = blk(c)
    with torch.no_grad():
        ms = timeit(step, reps)
        ms_b = timeit(lambda: rc.CorrBlock(f1, f2, L, r, build_mode=mode), reps)
        blk = rc.CorrBlock(f1, f2, L, r, build_mode=mode)
        ms_l = timeit(lambda: blk(coords[0]), reps * 4)
    bb, lb, *_ = alg_bytes(D, H, W, L, r)
    print(json.dumps({"config": name, "shape": [B, D, H, W, L, r], "iters": iters, "build_mode": mode, "ms_per_step": ms,
                      "pairs_per_s": B / (ms * 1e-3), "build_ms": ms_b, "build_gbs": B * bb / ms_b / 1e6, "build_frac_hbm": B * bb / ms_b / 1e6 / HBM,
                      "lookup_ms": ms_l, "lookup_gbs": B * lb / ms_l / 1e6, "lookup_frac_hbm": B * lb / ms_l / 1e6 / HBM}))


def train_config(name, B, D, H, W, L, r, iters, cls, kw, reps=5):
    f1, f2, coords = inputs(B, D, H, W, iters)
    ws = [torch.randn((B, L * (2 * r + 1) ** 2, H, W), device=dev) for _ in range(iters)]
    def step():
        a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
        blk = cls(a, b, L, r, **kw)
        if os.environ.get("TRAIN_LOSS_SUM") == "1":   # round-1 harness: 12 x (mul + sum) and their backward in torch
            loss = 0
            for c, w in zip(coords, ws): loss = loss + (blk(c) * w).sum()
            loss.backward()
        else:   # `out.backward(randn)` per iteration, accumulated -- no harness kernels in the timed step
            torch.autograd.backward([blk(c) for c in coords], ws)
    for _ in range(2): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):  # step latency: one step in flight at a time (median of per-step device times)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = _max_over_ranks(sorted(ts)[len(ts) // 2])
    print(json.dumps({"config": name, "shape": [B, D, H, W, L, r], "iters": iters, "class": cls.__name__, **kw, "ms_per_step_fwd_bwd": ms,
                      "harness": "loss = sum((out*w).sum()) in torch" if os.environ.get("TRAIN_LOSS_SUM") == "1" else "autograd.backward(outs, randn weights)",
                      "pairs_per_s": B / (ms * 1e-3)}))


def alt_config(name, B, D, H, W, L, r, reps=5, mode="tf32"):
    f1, f2, coords = inputs(B, D, H, W, 2)
    with torch.no_grad():
        blk = rc.AlternateCorrBlock(f1, f2, L, r, build_mode=mode)
        ms_ctor = timeit(lambda: rc.AlternateCorrBlock(f1, f2, L, r, build_mode=mode), reps)
        ms = timeit(lambda: blk(coords[1]), reps)
    *_, ab, af, _ = alg_bytes(D, H, W, L, r)
    print(json.dumps({"config": name, "mode": mode, "shape": [B, D, H, W, L, r], "ctor_ms": ms_ctor, "lookup_ms": ms,
                      "pair_iters_per_s": B / (ms * 1e-3), "alg_gbs": B * ab / ms / 1e6, "alg_tflops": B * af / ms / 1e9}))


which = sys.argv[1:] or ["1", "2", "3", "4", "5a", "5b"]
if "1" in which:
    fwd_config("1 RAFT-small 368x496 B1 12it", 1, 128, 46, 62, 4, 3, 12, "f16")
    fwd_config("1 RAFT-small 368x496 B1 12it", 1, 128, 46, 62, 4, 3, 12, "tf32")
    fwd_config("1 RAFT-small 368x496 B1 12it", 1, 128, 46, 62, 4, 3, 12, "fp32")
if "2" in which:
    fwd_config("2 RAFT-basic 436x1024 B8 32it", 8, 256, 55, 128, 4, 4, 32, "f16")
    fwd_config("2 RAFT-basic 436x1024 B8 32it", 8, 256, 55, 128, 4, 4, 32, "tf32")
    fwd_config("2 RAFT-basic 436x1024 B1 32it", 1, 256, 55, 128, 4, 4, 32, "f16")
if "3" in which:
    train_config("3 training 368x496 B10 12it fwd+bwd", 10, 256, 46, 62, 4, 4, 12, rc.CorrBlock, {"build_mode": "f16"}, reps=20)
    train_config("3 training (alternate) B2 2it fwd+bwd", 2, 256, 46, 62, 4, 4, 2, rc.AlternateCorrBlock, {}, reps=2)
if "4" in which:
    fwd_config("4 raft_uncertainty KITTI 375x1242 B16 24it", 16, 256, 47, 156, 4, 4, 24, "f16", reps=10)
    fwd_config("4 raft_uncertainty KITTI 375x1242 B16 24it", 16, 256, 47, 156, 4, 4, 24, "tf32", reps=10)
if "5a" in which:
    alt_config("5a alt 1080x1920 B4", 4, 256, 135, 240, 4, 4, mode="tf32")
    if WORLD == 1:
        alt_config("5a alt 1080x1920 B4", 4, 256, 135, 240, 4, 4, mode="fp32")
if "5b" in which:
    alt_config("5b alt 2160x3840 B4", 4, 256, 270, 480, 4, 4, reps=3, mode="tf32")
    if WORLD == 1:
        alt_config("5b alt 2160x3840 B4", 4, 256, 270, 480, 4, 4, reps=3, mode="fp32")
if WORLD > 1:
    dist.destroy_process_group()
