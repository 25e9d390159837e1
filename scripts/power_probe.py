"""Clocks and board power while ONE kernel family runs back to back for a few seconds (nvidia-smi at 20 ms):
is the build / lookup power-capped?  usage: power_probe.py build|lookup|head [seconds]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optical_flow_dnn_b200 as rc
what = sys.argv[1] if len(sys.argv) > 1 else "build"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
dev = torch.device("cuda:0")
B, D, H, W = 8, int(os.environ.get("D", "256")), 55, 128
f1 = torch.randn((B, D, H, W), device=dev); f2 = torch.randn((B, D, H, W), device=dev)
ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
coords = (torch.stack([xs, ys]).float()[None].expand(B, 2, H, W) + 3 * torch.randn((B, 2, H, W), device=dev)).contiguous()
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown",
                         "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
with torch.no_grad():
    blk = rc.CorrBlock(f1, f2, 4, 4)
    torch.cuda.synchronize(); time.sleep(0.5)
    t0 = time.time(); per = []
    while time.time() - t0 < secs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            if what == "build": blk = rc.CorrBlock(f1, f2, 4, 4)
            else: out = blk(coords)
        e1.record(); torch.cuda.synchronize()
        per.append((time.time() - t0, e0.elapsed_time(e1) / 50 * 1e3))
    t1 = time.time()
time.sleep(0.3); proc.terminate()
print(f"{what}: us per launch over time:", " ".join(f"{t:.1f}s:{u:.0f}" for t, u in per[:: max(1, len(per) // 12)]))
busy = [l for (ts, l) in lines if t0 + 0.2 < ts < t1]
idle = [l for (ts, l) in lines if ts < t0 - 0.1]
def col(ls, i):
    v = []
    for l in ls:
        try: v.append(float(l.split(",")[i]))
        except Exception: pass
    return v
for name, ls in (("idle", idle), ("busy", busy)):
    sm, pw = col(ls, 0), col(ls, 2)
    if sm:
        cap = sum("Active" in l.split(",")[4] for l in ls)
        print(f"  {name}: {len(sm)} samples, SM MHz min/median/max {min(sm):.0f}/{sorted(sm)[len(sm)//2]:.0f}/{max(sm):.0f}, power W median/max {sorted(pw)[len(pw)//2]:.0f}/{max(pw):.0f}, sw_power_cap active in {cap}")
