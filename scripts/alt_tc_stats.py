"""Tensor-core on-the-fly lookup: lattice tiles per source block (from the bbox work space), time, redundancy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
def run(tag, B, D, H, W, L, r, sigma, smooth, reps=5):
    g = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn((B, D, H, W), generator=g, device=dev); f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    grid = torch.stack([xs, ys]).float()[None].expand(B, 2, H, W)
    flow = sigma * torch.randn((B, 2, H, W), generator=g, device=dev)
    if smooth:
        flow = torch.nn.functional.avg_pool2d(flow, 5, 1, 2, count_include_pad=False) * 5
    coords = (grid + flow).contiguous()
    with torch.no_grad():
        blk = rc.AlternateCorrBlock(f1, f2, L, r, build_mode="tf32")
        for _ in range(2): blk(coords)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): blk(coords)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bb = blk._ws.view(torch.int32).view(-1, L, 4)
    z = bb[..., 2]
    cnt = (z & 0xffff) * bb[..., 3]
    tw, th = 32 >> ((z >> 16) & 3), (z >> 20) & 63   # lattice tile width / height (bbox.z encoding of altcorr_tc.cu)
    tiles = cnt.sum(0).tolist()
    by_w = {w: int((cnt * (tw == w)).sum()) for w in (32, 16, 8)}
    pts = cnt * tw * th
    nblk = bb.shape[0]
    K = 2 * r + 1
    useful = B * H * W * L * (K + 1) ** 2
    tile_n = 256
    done = int(pts.sum()) * 128
    print(f"{tag}: {ms*1e3:8.1f} us  blocks={nblk} lattice tiles/level={tiles} (tiles by width: {by_w}, mean N {int(pts.sum()) / max(sum(tiles), 1):.0f}) per block={sum(tiles)/nblk:.1f} "
          f"redundancy={done/useful:.2f}x  MMA TFLOP/s={done*2*D/ms/1e9:.0f}  us/tile={ms*1e3*148/max(sum(tiles),1):.2f}")
which = sys.argv[1:] or ["a"]
for w in which:
    if w == "a":
        run("1080p smooth s4 ", 4, 256, 135, 240, 4, 4, 4.0, True)
        run("1080p zero flow ", 4, 256, 135, 240, 4, 4, 0.0, False)
        run("1080p noise s1  ", 4, 256, 135, 240, 4, 4, 1.0, False)
        run("1080p noise s4  ", 4, 256, 135, 240, 4, 4, 4.0, False)
        run("1080p D128 small", 4, 128, 135, 240, 4, 3, 4.0, True)
    if w == "c":
        run("1080p r3 smooth ", 4, 256, 135, 240, 4, 3, 4.0, True)
        run("1080p r3 zero   ", 4, 256, 135, 240, 4, 3, 0.0, False)
    if w == "b":
        run("2160p smooth s4 ", 4, 256, 270, 480, 4, 4, 4.0, True, reps=3)
        run("2160p noise s4  ", 4, 256, 270, 480, 4, 4, 4.0, False, reps=2)
