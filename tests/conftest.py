import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_CASES = ["small_r3_odd", "basic_r4", "levels3_r4"]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", f"corr_{request.param}.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(params=["small", "basic"])
def golden_convc1(request):
    """Reference CorrBlock + the reference motion encoder's convc1 + ReLU (tests/golden/make_golden_convc1.py)."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", f"convc1_{request.param}.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(params=["flow", "flow_var"])
def golden_upsample(request):
    """The reference's own RAFT.upsample_flow and its autograd gradients (tests/golden/make_golden_upsample.py)."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", f"upsample_{request.param}.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(params=["basic", "small"])
def golden_fnet_head(request):
    """The reference encoders' own conv2 head + the reference CorrBlock on its output (tests/golden/make_golden_fnet_head.py)."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", f"fnet_head_{request.param}.npz"))
    return {k: g[k] for k in g.files}
