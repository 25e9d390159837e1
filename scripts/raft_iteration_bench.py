"""What the correlation-side kernels buy inside a RAFT-basic refinement loop (raft/core/raft.py:141-176).

The update block (BasicMotionEncoder -> SepConvGRU -> FlowHead + mask head, raft/core/update.py:153-273, written here
from its architecture with random weights; cuDNN convolutions in both arms) is the CALLER of this library and is the
same in both arms.  What differs is everything around it in the loop:

  stock : the reference recipe in stock PyTorch -- matmul / avg_pool2d / grid_sample CorrBlock, F.relu(convc1(corr)),
          softmax/unfold/sum upsample_flow
  ours  : CorrBlock (tcgen05 build, TMA lookup), CorrBlock.lookup_convc1 (lookup fused with convc1 + ReLU),
          upsample_flow (one pass)

Prints per-arm ms for build + `iters` refinement iterations, and the final-flow difference between the arms."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import optical_flow_dnn_b200 as rc
from stock_torch_gpu import StockCorrBlock

dev = torch.device("cuda:0")


class UpdateBlock(nn.Module):
    def __init__(self, cor_planes=324, hidden=128):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(256, 126, 3, padding=1)
        cin = hidden + 128 + hidden
        self.gz1, self.gr1, self.gq1 = (nn.Conv2d(cin, hidden, (1, 5), padding=(0, 2)) for _ in range(3))
        self.gz2, self.gr2, self.gq2 = (nn.Conv2d(cin, hidden, (5, 1), padding=(2, 0)) for _ in range(3))
        self.fh1 = nn.Conv2d(hidden, 256, 3, padding=1)
        self.fh2 = nn.Conv2d(256, 2, 3, padding=1)
        self.m1 = nn.Conv2d(128, 256, 3, padding=1)
        self.m2 = nn.Conv2d(256, 576, 1)

    def forward(self, net, inp, corr, flow):
        if isinstance(corr, tuple):                      # (corr_fn, coords)
            cor = corr[0].lookup_convc1(corr[1], self.convc1.weight, self.convc1.bias)
        else:
            cor = F.relu(self.convc1(corr))
        cor = F.relu(self.convc2(cor))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        mot = torch.cat([F.relu(self.conv(torch.cat([cor, flo], 1))), flow], 1)
        x = torch.cat([inp, mot], 1)
        for gz, gr, gq in ((self.gz1, self.gr1, self.gq1), (self.gz2, self.gr2, self.gq2)):
            hx = torch.cat([net, x], 1)
            z, r = torch.sigmoid(gz(hx)), torch.sigmoid(gr(hx))
            q = torch.tanh(gq(torch.cat([r * net, x], 1)))
            net = (1 - z) * net + z * q
        return net, self.m2(F.relu(self.m1(net))), self.fh2(F.relu(self.fh1(net)))


def stock_upsample(flow, mask):
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    return torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W)


def refine(arm, op, f1, f2, net, inp, iters):
    B, _, H, W = f1.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    c0 = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1)
    c1 = c0.clone()
    corr_fn = StockCorrBlock(f1, f2, 4, 4) if arm == "stock" else rc.CorrBlock(f1, f2, 4, 4)
    up = None
    for _ in range(iters):
        corr = corr_fn(c1) if arm == "stock" else (corr_fn, c1)
        net, mask, delta = op(net, inp, corr, c1 - c0)
        c1 = c1 + 0.05 * delta          # random weights: keep the synthetic flow from running away
        up = stock_upsample(c1 - c0, mask) if arm == "stock" else rc.upsample_flow((c1 - c0).contiguous(), mask)
    return up


def main():
    torch.manual_seed(0)
    op = UpdateBlock().to(dev).eval()
    res = []
    for (B, iters) in ((1, 32), (8, 12)):
        H, W, D = 55, 128, 256
        g = torch.Generator(device=dev).manual_seed(5)
        base = F.avg_pool2d(torch.randn((B, D, H + 8, W + 8), generator=g, device=dev), 3, 1, 1) * 3.0
        f1 = base[:, :, 4:4 + H, 4:4 + W].contiguous()
        f2 = (base[:, :, 3:3 + H, 6:6 + W] + 0.1 * torch.randn((B, D, H, W), generator=g, device=dev)).contiguous()
        net = torch.tanh(torch.randn((B, 128, H, W), generator=g, device=dev))
        inp = torch.relu(torch.randn((B, 128, H, W), generator=g, device=dev))
        out = {}
        with torch.no_grad():
            for arm in ("stock", "ours"):
                for _ in range(2): up = refine(arm, op, f1, f2, net, inp, iters)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5): up = refine(arm, op, f1, f2, net, inp, iters)
                e1.record(); torch.cuda.synchronize()
                out[arm] = (e0.elapsed_time(e1) / 5, up)
        epe = float((out["stock"][1] - out["ours"][1]).pow(2).sum(1).sqrt().mean())
        res.append({"B": B, "iters": iters, "stock_ms": out["stock"][0], "ours_ms": out["ours"][0],
                    "speedup": out["stock"][0] / out["ours"][0], "final_flow_epe_between_arms_px": epe})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
