"""Random-shape fuzz on the GPU: every tensor-core path against the exact-fp32 build of this library (itself pinned
against the reference goldens), odd sizes, tails, small feature dims.  Prints the worst error per path and any failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optical_flow_dnn_b200 as rc
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
worst, fails = {}, []

def rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))

def note(path, e, tol, case):
    worst[path] = max(worst.get(path, 0.0), e)
    if not (e <= tol):
        fails.append((path, e, case))

for case_i in range(n_cases):
    B = int(rng.integers(1, 4)); D = int(rng.choice([8, 24, 32, 48, 64, 96, 128, 160, 256, 320]))
    H = int(rng.integers(8, 41)); W = int(rng.integers(8, 71)); r = int(rng.integers(1, 5))
    L = int(rng.integers(1, 5))
    while (H >> (L - 1)) < 2 or (W >> (L - 1)) < 2: L -= 1
    case = (B, D, H, W, L, r)
    g = torch.Generator(device=dev).manual_seed(1000 + case_i)
    mag = float(10.0 ** rng.integers(-6, 7))
    f1 = torch.randn((B, D, H, W), generator=g, device=dev) * mag
    f2 = torch.randn((B, D, H, W), generator=g, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    coords = (torch.stack([xs, ys]).float()[None].expand(B, 2, H, W) + 5 * torch.randn((B, 2, H, W), generator=g, device=dev)).contiguous()
    try:
        with torch.no_grad():
            ref_blk = rc.CorrBlock(f1, f2, L, r, build_mode="fp32")
            ref = ref_blk(coords)
            for mode, tol in (("f16", 1e-3), ("tf32", 1e-3), ("f16x3", 2e-5)):
                blk = rc.CorrBlock(f1, f2, L, r, build_mode=mode)
                note(f"CorrBlock[{mode}] lookup", rel(blk(coords), ref), tol, case)
                for l, (a, c) in enumerate(zip(blk.corr_pyramid, ref_blk.corr_pyramid)):
                    note(f"CorrBlock[{mode}] level", rel(a, c), tol, case + (l,))
            if D % 4 == 0:
                note("AlternateCorrBlock (default)", rel(rc.AlternateCorrBlock(f1, f2, L, r)(coords), ref), 1e-3, case)
            cout = int(rng.choice([16, 48, 96, 128, 256]))
            K = 2 * r + 1
            w = torch.randn((cout, L * K * K), generator=g, device=dev) / (L * K * K) ** 0.5
            bias = torch.randn((cout,), generator=g, device=dev) * float(ref.abs().max())
            want = torch.relu(torch.einsum("oc,bchw->bohw", w.double(), ref.double()) + bias.double().view(1, -1, 1, 1))
            got = rc.CorrBlock(f1, f2, L, r).lookup_convc1(coords, w, bias)
            note("lookup_convc1", float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30)), 2e-3, case + (cout,))
            Cin = int(rng.choice([8, 32, 96, 100, 128]))
            x = torch.randn((2 * B, Cin, H, W), generator=g, device=dev) * mag
            hw = torch.randn((D, Cin), generator=g, device=dev) / Cin ** 0.5
            hb = torch.randn((D,), generator=g, device=dev) * mag
            f = torch.einsum("dc,bchw->bdhw", hw.double(), x.double()) + hb.double().view(1, -1, 1, 1)
            href = rc.CorrBlock(f[:B].float(), f[B:].float(), L, r, build_mode="fp32")(coords)
            note("from_encoder_head", rel(rc.CorrBlock.from_encoder_head(x, hw, hb, num_levels=L, radius=r)(coords), href), 1.5e-3, case + (Cin,))
            if D % 32 == 0:
                note("Alternate.from_encoder_head", rel(rc.AlternateCorrBlock.from_encoder_head(x, hw, hb, num_levels=L, radius=r)(coords), href), 1.5e-3, case + (Cin,))
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        fails.append(("EXCEPTION", repr(e)[:200], case))
        break
for k, v in sorted(worst.items()): print(f"{k:32s} worst rel-to-max error {v:.2e}")
print(f"{n_cases} cases, {len(fails)} failures")
for f in fails[:20]: print("FAIL", f)
