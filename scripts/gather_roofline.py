"""What the memory system charges for the lookup's box list, measured: the production kernel with
RAFTCORR_LOOKUP_EXP=1 issues exactly the same TMA boxes and waits for them but evaluates and writes nothing ("gather
peak"); EXP=2 evaluates and writes from stale shared memory without any box.  Bench shape (B=8, 55x128, D=256, r=4, L=4),
CUDA graph of 12 lookups with the bench's coordinates.  Prints one JSON line.

    python scripts/gather_roofline.py            # runs itself three times (the switch is read per process)
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one():
    sys.path.insert(0, ROOT)
    import torch
    import optical_flow_dnn_b200 as rc
    dev = torch.device("cuda:0")
    B, D, H, W = 8, 256, 55, 128
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = torch.nn.functional.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, 1, 2, count_include_pad=False) * 5
    coords = [(grid + flow).contiguous()]
    for _ in range(11): coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
    with torch.no_grad():
        blk = rc.CorrBlock(f1, f2, 4, 4)
        for c in coords: blk(c)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for c in coords: o = blk(c)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(7):
            e0.record()
            for _ in range(5): gr.replay()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 60 * 1e3)
    print(json.dumps({"exp": os.environ.get("RAFTCORR_LOOKUP_EXP", "0"), "us_per_launch": sorted(ts)[3]}))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
        sys.exit(0)
    res = {}
    for exp in ("0", "1", "2"):
        env = dict(os.environ, RAFTCORR_LOOKUP_EXP=exp)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
        res[exp] = json.loads(out.stdout.strip().splitlines()[-1])["us_per_launch"]
    alg = 163553280  # algorithmic bytes per launch (8 pairs x 20.44 MB)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"lookup_us": res["0"], "gather_only_us": res["1"], "evaluate_and_store_only_us": res["2"],
                      "algorithmic_bytes_per_launch": alg, "frac_of_hbm_peak": alg / (res["0"] * 1e-6) / 1e9 / peak,
                      "gather_only_frac_of_lookup": res["1"] / res["0"],
                      "what": "B=8, 55x128, r=4, L=4; CUDA graph of 12 lookups; gather only = same TMA boxes and waits, no evaluation, no output"}))
