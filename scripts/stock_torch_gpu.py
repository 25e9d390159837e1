"""Like-for-like GPU baseline: the reference CorrBlock's own recipe -- torch.matmul, F.avg_pool2d,
F.grid_sample, cat + permute (raft/core/corr.py:18-116, utils.py:57-85) -- run with stock PyTorch CUDA kernels on the
same B200, next to this library, on the bench workload (B=8, 55x128, D=256, r=4, build + 12 lookups).
Self-contained on purpose (scripts/ must not import oracle/); checked against the library's output before timing."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import optical_flow_dnn_b200 as rc


class StockCorrBlock:
    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        B, D, H, W = fmap1.shape
        self.num_levels, self.radius = num_levels, radius
        corr = torch.matmul(fmap1.view(B, D, H * W).transpose(1, 2), fmap2.view(B, D, H * W))
        corr = corr.view(B, H, W, 1, H, W) / math.sqrt(D)
        corr = corr.reshape(B * H * W, 1, H, W)
        self.pyr = [corr]
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, 2, stride=2)
            self.pyr.append(corr)

    def __call__(self, coords):
        r = self.radius
        coords = coords.permute(0, 2, 3, 1)
        B, H, W, _ = coords.shape
        out = []
        d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
        for i, corr in enumerate(self.pyr):
            c = coords.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta
            h, w = corr.shape[-2:]
            gx = 2 * c[..., 0] / (w - 1) - 1
            gy = 2 * c[..., 1] / (h - 1) - 1
            s = F.grid_sample(corr, torch.stack([gx, gy], dim=-1), align_corners=True)
            out.append(s.view(B, H, W, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def main():
    dev = torch.device("cuda:0")
    B, D, H, W, L, r, iters = 8, 256, 55, 128, 4, 4, 12
    g = torch.Generator(device=dev).manual_seed(1234)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev)
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = F.avg_pool2d(4.0 * torch.randn((B, 2, H, W), generator=g, device=dev), 5, stride=1, padding=2,
                        count_include_pad=False) * 5.0
    coords = [(grid + flow).contiguous()]
    for _ in range(iters - 1):
        coords.append((coords[-1] + 0.5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous())
    res = {}
    with torch.no_grad():
        ours = rc.CorrBlock(f1, f2, L, r)(coords[0])
        for name, tf32 in (("stock_fp32", False), ("stock_tf32_matmul", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ref = StockCorrBlock(f1, f2, L, r)(coords[0])
            res[name + "_vs_ours_rel"] = float((ref - ours).abs().max() / ref.abs().max())

            def step():
                blk = StockCorrBlock(f1, f2, L, r)
                for c in coords:
                    o = blk(c)
                return o
            for _ in range(3): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            # build alone
            e0.record()
            for _ in range(10): blk = StockCorrBlock(f1, f2, L, r)
            e1.record(); torch.cuda.synchronize()
            bms = e0.elapsed_time(e1) / 10
            res[name] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "build_ms": bms, "lookup_ms": (ms - bms) / iters}
        def step_ours():
            blk = rc.CorrBlock(f1, f2, L, r)
            for c in coords:
                o = blk(c)
            return o
        for _ in range(3): step_ours()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step_ours()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        res["ours_tf32"] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3}
    res["workload"] = f"B={B} pairs, {H}x{W}, D={D}, r={r}, L={L}, build + {iters} lookups; torch {torch.__version__}"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
