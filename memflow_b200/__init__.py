"""memflow_b200 -- B200-native (sm_100a) implementation of MEMFlow's batched phase-space mapping and
importance-sampling hot path, behind the reference's own Python API.

    from memflow_b200.phasespace.phasespace import PhaseSpace          # memflow.phasespace.phasespace
    from memflow_b200.transforms import MonotonicRQSTransform          # zuko.transforms
    from memflow_b200.estimator import accumulate_weights, mc_estimate, ImportanceSampler

All compute goes through hand-written CUDA kernels (memflow_b200/csrc) via the C ABI of
include/memflow_b200.h. There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from ._lib import MemflowB200Error  # noqa: F401
