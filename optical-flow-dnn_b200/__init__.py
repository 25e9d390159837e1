"""b200-raft-corr: B200-native (sm_100a) RAFT correlation hot path.

Drop-in for ``raft/core/corr.py`` of andrei-iosif/optical-flow-dnn::

    from optical_flow_dnn_b200 import CorrBlock, AlternateCorrBlock

Host code is Python/PyTorch (device memory, streams, autograd); all arithmetic runs in
hand-written CUDA behind the C ABI of ``include/raftcorr.h`` (``libraftcorr_b200.so``).
"""
from .corr import AlternateCorrBlock, CorrBlock, set_default_build_mode  # noqa: F401
from ._lib import RaftCorrError  # noqa: F401
from .upsample import upsample_flow  # noqa: F401

__all__ = ["CorrBlock", "AlternateCorrBlock", "RaftCorrError", "set_default_build_mode", "upsample_flow"]
