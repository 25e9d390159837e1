"""Imports the reference's own caller code (``raft/core/raft.py`` etc.) from the staged, byte-identical copies under
``baseline/_ref/`` (scripts/vendor_reference.py) -- test infrastructure, never imported by the package.

Two arms share the top-level module name ``core`` (raft.py:5-8 imports ``core.update``, ``core.extractor``,
``core.corr``, ``core.utils.utils``), so each arm is imported with a clean ``sys.modules`` slate and its modules are
returned as a namespace:

    stock  -- the reference's own ``core/corr.py`` (pure PyTorch CorrBlock; AlternateCorrBlock over the reference's own
              compiled ``alt_cuda_corr`` extension)
    dropin -- identical files, ``core/corr.py`` replaced by the INTEGRATION.md 2a re-export of this library
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, arm, "raft", "core", "raft.py")) for arm in ("stock", "dropin"))


def reference_alt_extension_available() -> bool:
    import glob

    return bool(glob.glob(os.path.join(REF_DIR, "stock", "alt_cuda_corr*.so")))


def _purge():
    for name in [n for n in sys.modules if n == "core" or n.startswith("core.") or n == "alt_cuda_corr"]:
        del sys.modules[name]


def load_arm(arm: str):
    """Returns the arm's ``core.raft`` module (``.RAFT``, ``.CorrBlock``, ``.AlternateCorrBlock``)."""
    assert arm in ("stock", "dropin")
    base = os.path.join(REF_DIR, arm)
    paths = [os.path.join(base, "raft")] + ([base] if arm == "stock" else [])  # stock: its own alt_cuda_corr*.so
    _purge()
    sys.path[:0] = paths
    try:
        mod = importlib.import_module("core.raft")
        corr_mod = sys.modules["core.corr"]
        utils_mod = sys.modules["core.utils.utils"]
    finally:
        for p in paths:
            sys.path.remove(p)
        _purge()
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(base)), mod.__file__
    mod._corr_module = corr_mod
    mod._utils_module = utils_mod  # InputPadder (utils.py:7-27), coords_grid, bilinear_sampler
    return mod


def raft_args(small=False, uncertainty=False, alternate_corr=False, mixed_precision=False, residual_variance=False,
              log_variance=False, dropout=0.0):
    """The Namespace train.py / evaluate.py build with argparse (raft/train.py:212-250, raft/core/raft.py:24-45)."""
    return argparse.Namespace(small=small, uncertainty=uncertainty, alternate_corr=alternate_corr,
                              mixed_precision=mixed_precision, residual_variance=residual_variance,
                              log_variance=log_variance, dropout=dropout)
