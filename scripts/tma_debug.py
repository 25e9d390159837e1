import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import optical_flow_dnn_b200 as rc
from oracle import corr_oracle as O
dev = torch.device("cuda:0")
B, D, H, W, L, r = (int(x) for x in (sys.argv[1:7] if len(sys.argv) > 6 else (1, 32, 16, 24, 2, 4)))
rng = np.random.default_rng(0)
f1 = rng.normal(size=(B, D, H, W)).astype(np.float32); f2 = rng.normal(size=(B, D, H, W)).astype(np.float32)
co = O.coords_grid(B, H, W) + rng.normal(0, 2, (B, 2, H, W)).astype(np.float32)
blk = rc.CorrBlock(torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev), L, r, build_mode="fp32")
torch.cuda.synchronize()
t0 = time.time()
try:
    out = blk(torch.from_numpy(co).to(dev)); torch.cuda.synchronize()
    ref = O.lookup(O.corr_pyramid(f1, f2, L, np.float64), co, r)
    print("ok in %.3fs, err %.2e" % (time.time() - t0, np.abs(out.cpu().numpy() - ref).max()))
except Exception as e:
    print("FAILED after %.3fs: %s" % (time.time() - t0, str(e)[:200]))
