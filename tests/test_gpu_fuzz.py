"""Random-shape sweep on the GPU (scripts/fuzz_shapes.py): odd sizes, pixel/channel tails, feature dims 8..320, input
magnitudes 1e-6..1e6 -- every tensor-core path (f16 / tf32 builds, on-the-fly class, fused lookup+convc1, fused encoder
head for both classes) against this library's exact-fp32 build, which the golden tests pin to the reference."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [3, 4])
def test_random_shapes(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_shapes.py"), str(seed), "30"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "30 cases, 0 failures" in r.stdout, r.stdout[-3000:]
