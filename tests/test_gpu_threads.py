"""Host-side re-entrancy: several Python threads drive ONE device at the same time, each on its own stream.

The reference is called from nn.DataParallel worker threads (raft/train.py:71, one thread per device; covered by
test_two_devices_from_one_process); a serving process may also run several requests on one GPU.  The C ABI promises
"no global mutable state, re-entrant" (include/raftcorr.h): the forward paths are deterministic, so every thread must
reproduce, bit for bit, what the same inputs gave when run alone.
"""
import threading

import pytest
import torch

from oracle import corr_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [(2, 64, 24, 40, 4, 4), (1, 128, 46, 62, 4, 3), (2, 64, 23, 37, 3, 4), (1, 96, 31, 45, 4, 2)]


def _case(i):
    B, D, H, W, L, r = SHAPES[i % len(SHAPES)]
    d = torch.device("cuda:0")
    g = torch.Generator(device=d).manual_seed(100 + i)
    f1 = torch.randn((B, D, H, W), generator=g, device=d)
    f2 = torch.randn((B, D, H, W), generator=g, device=d)
    coords = torch.from_numpy(O.coords_grid(B, H, W)).to(d) + 2.5 * torch.randn((B, 2, H, W), generator=g, device=d)
    cw = torch.randn((64, L * (2 * r + 1) ** 2), generator=g, device=d) / 18.0
    flow = torch.randn((B, 2, H, W), generator=g, device=d)
    mask = torch.randn((B, 576, H, W), generator=g, device=d)
    return (L, r), (f1, f2, coords, cw, flow, mask)


def _run(pkg, cfg, tensors, mode):
    (L, r), (f1, f2, coords, cw, flow, mask) = cfg, tensors
    with torch.no_grad():
        blk = pkg.CorrBlock(f1, f2, num_levels=L, radius=r, build_mode=mode)
        outs = [blk(coords + 0.25 * k) for k in range(3)]
        outs.append(blk.lookup_convc1(coords, cw, None))
        outs.append(pkg.AlternateCorrBlock(f1, f2, num_levels=L, radius=r)(coords))
        outs.append(pkg.upsample_flow(flow, mask))
    return outs


def test_threads_on_one_device_reproduce_the_serial_results():
    import optical_flow_dnn_b200 as pkg

    n_threads, rounds = 6, 4
    modes = ["f16", "tf32", "fp32", "f16x3"]
    cases = [_case(i) for i in range(n_threads)]
    serial = [[t.clone() for t in _run(pkg, c, x, modes[i % 4])] for i, (c, x) in enumerate(cases)]
    torch.cuda.synchronize()
    errors, start = [], threading.Barrier(n_threads)

    def work(i):
        try:
            cfg, tensors = cases[i]
            s = torch.cuda.Stream(device="cuda:0")
            start.wait()
            with torch.cuda.stream(s):
                for _ in range(rounds):
                    outs = _run(pkg, cfg, tensors, modes[i % 4])
                    s.synchronize()
                    for k, (a, b) in enumerate(zip(outs, serial[i])):
                        if not torch.equal(a, b):
                            errors.append((i, k, float((a - b).abs().max())))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
