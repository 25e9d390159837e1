"""CPU-side checks: the C-ABI library loads and exports every symbol include/raftcorr.h declares,
host-only entry points behave, and the Python mirror refuses to run without CUDA tensors."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from optical_flow_dnn_b200 import _lib

    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "raftcorr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(raftcorr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in raftcorr.h but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(lib.SIGNATURES) == names


def test_library_has_no_python_or_torch_dependency(lib):
    import subprocess

    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


def test_layout_matches_oracle_shapes(lib):
    from oracle.corr_oracle import level_shapes

    for (B, H, W, L) in [(8, 55, 128, 4), (1, 46, 62, 4), (16, 47, 156, 4), (2, 9, 12, 3)]:
        for il in (0, 1, 2):  # plain rows / row pairs / row quads interleaved (the last group is padded with zero rows)
            lay = lib.make_layout(B, H, W, L, interleaved=il)
            assert lay.interleaved == il
            assert [(lay.h[l], lay.w[l]) for l in range(L)] == level_shapes(H, W, L)
            off = 0
            for l in range(L):
                assert lay.pitch[l] % 8 == 0 and lay.pitch[l] >= lay.w[l]
                assert lay.offset[l] == off and off % 64 == 0
                rg = (1, 2, 4)[il]
                rows = (lay.h[l] + rg - 1) // rg * rg
                off += -(-(B * H * W * rows * lay.pitch[l]) // 64) * 64
            assert lay.total_floats == off
        assert lib.make_layout(B, H, W, L, 2).total_floats >= lib.make_layout(B, H, W, L, 1).total_floats >= lib.make_layout(B, H, W, L, 0).total_floats


def test_layout_rejects_collapsed_levels_and_bad_counts(lib):
    with pytest.raises(lib.RaftCorrError, match="divides by"):
        lib.make_layout(1, 8, 8, 4)
    with pytest.raises(lib.RaftCorrError):
        lib.make_layout(1, 64, 64, 0)
    with pytest.raises(lib.RaftCorrError):
        lib.make_layout(1, 64, 64, 9)


def test_null_pointers_are_rejected_without_a_gpu(lib):
    L = lib.lib()
    lay = lib.make_layout(1, 16, 16, 2)
    rc = L.raftcorr_lookup_fwd(None, ctypes.byref(lay), None, 4, None, 0, None)
    assert rc != 0 and b"NULL" in L.raftcorr_last_error()
    rc = L.raftcorr_build_fwd(None, None, 16, None, ctypes.byref(lay), 0, None, 0, 0, None)
    assert rc != 0


def test_batched_lookup_adjoint_host_side_checks(lib):
    """raftcorr_lookup_bwd_batched: the shared-memory fit query and the argument checks run without a GPU."""
    L = lib.lib()
    assert L.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lib.make_layout(8, 55, 128, 4, 2)), 4) == 1   # config 2
    assert L.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lib.make_layout(10, 46, 62, 4, 2)), 4) == 1   # config 3
    assert L.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lib.make_layout(1, 135, 240, 4, 2)), 4) == 0  # 1080p: per-iteration path
    assert L.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lib.make_layout(1, 64, 64, 5, 0)), 4) == 0    # > 4 levels
    assert L.raftcorr_lookup_bwd_batched_supported(ctypes.byref(lib.make_layout(1, 32, 32, 2, 0)), 5) == 0    # radius
    lay = lib.make_layout(1, 16, 16, 2)
    rc = L.raftcorr_lookup_bwd_batched(None, None, 1, None, ctypes.byref(lay), 4, 0, 0, 0, None)
    assert rc != 0 and b"NULL" in L.raftcorr_last_error()
    assert L.raftcorr_build_bwd_wants_tf32(256, lib.BUILD_F16_TC) == 1
    assert L.raftcorr_build_bwd_wants_tf32(256, lib.BUILD_FP32_SIMT) == 0
    assert L.raftcorr_build_bwd_wants_tf32(320, lib.BUILD_TF32_TC) == 0  # D > 256: CUDA-core products


def test_cpu_tensors_fail_loudly(lib):
    import torch

    import optical_flow_dnn_b200 as m

    f = torch.zeros(1, 8, 16, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.CorrBlock(f, f)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.AlternateCorrBlock(f, f)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optical-flow-dnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"
