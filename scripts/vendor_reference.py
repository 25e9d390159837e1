"""Stage the UNMODIFIED reference (andrei-iosif/optical-flow-dnn) under the git-ignored ``baseline/_ref/`` so that it
travels to the GPU box with the repo snapshot (``/root/reference`` does not exist there).  SURVEY.md 7.1 step 1.

    baseline/_ref/stock/raft/core/{raft,update,extractor,corr}.py, utils/utils.py   -- byte-identical copies
    baseline/_ref/dropin/raft/core/{raft,update,extractor}.py, utils/utils.py       -- byte-identical copies
    baseline/_ref/dropin/raft/core/corr.py                                          -- the INTEGRATION.md 2a re-export
    baseline/_ref/stock/alt_cuda_corr*.so    -- the reference's own extension (raft/alt_cuda_corr), built from its
                                                sources where they lie, arch sm_100 (TORCH_CUDA_ARCH_LIST=10.0)

Nothing here is product code and nothing in the package imports it: tests (`tests/test_gpu_raft_reference.py`), the
`gpu_baseline` / `--impl reference` legs of bench.py and scripts/ use it as the checker / baseline.  Called by
``__graft_entry__.build()`` when ``/root/reference`` is present; a no-op otherwise (the GPU box uses the staged files).
"""
from __future__ import annotations

import filecmp
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RAFTCORR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
CALLER_FILES = ["core/__init__.py", "core/raft.py", "core/update.py", "core/extractor.py", "core/utils/__init__.py",
                "core/utils/utils.py"]
SHIM = '''# raft/core/corr.py -- INTEGRATION.md 2a: the only file a maintainer changes
import os
import sys

# this copy lives at <repo>/baseline/_ref/dropin/raft/core/corr.py; a maintainer writes the checkout path here
sys.path.insert(0, os.environ.get("RAFTCORR_B200_ROOT", os.path.abspath(os.path.join(os.path.dirname(__file__), *[os.pardir] * 5))))
from optical_flow_dnn_b200 import CorrBlock, AlternateCorrBlock  # noqa: E402,F401  same ctor / __call__ / attributes
'''


def _copy(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
        shutil.copyfile(src, dst)


def stage_python() -> bool:
    src_root = os.path.join(REF, "raft")
    if not os.path.isdir(src_root):
        return False
    for arm in ("stock", "dropin"):
        for rel in CALLER_FILES:
            _copy(os.path.join(src_root, rel), os.path.join(DST, arm, "raft", rel))
    _copy(os.path.join(src_root, "core/corr.py"), os.path.join(DST, "stock", "raft", "core/corr.py"))
    shim = os.path.join(DST, "dropin", "raft", "core/corr.py")
    text = SHIM
    if not os.path.exists(shim) or open(shim).read() != text:
        with open(shim, "w") as f:
            f.write(text)
    return True


def build_alt_cuda_corr(force: bool = False) -> str | None:
    """Build the reference's own ``alt_cuda_corr`` torch extension (raft/alt_cuda_corr/{correlation.cpp,
    correlation_kernel.cu}, its own setup.py) for sm_100 from a scratch copy (the reference tree is read-only and its
    setup.py writes next to the sources).  Returns the staged .so path, or None when it cannot be built here."""
    out_dir = os.path.join(DST, "stock")
    have = glob.glob(os.path.join(out_dir, "alt_cuda_corr*.so"))
    if have and not force:
        return have[0]
    src = os.path.join(REF, "raft", "alt_cuda_corr")
    if not os.path.isdir(src):
        return None
    tmp = tempfile.mkdtemp(prefix="alt_cuda_corr_")
    try:
        work = os.path.join(tmp, "alt_cuda_corr")
        shutil.copytree(src, work)
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="4")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=work, env=env, capture_output=True,
                           text=True)
        built = glob.glob(os.path.join(work, "alt_cuda_corr*.so"))
        if r.returncode != 0 or not built:
            sys.stderr.write("reference alt_cuda_corr did not build here:\n" + r.stdout[-2000:] + r.stderr[-4000:] + "\n")
            return None
        os.makedirs(out_dir, exist_ok=True)
        dst = os.path.join(out_dir, os.path.basename(built[0]))
        shutil.copyfile(built[0], dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main(force: bool = False):
    ok = stage_python()
    so = build_alt_cuda_corr(force) if ok else None
    return ok, so


if __name__ == "__main__":
    print(main(force="--force" in sys.argv))
