"""computervision.pytorch_b200 — B200-native (sm_100a) detection post-processing.

The dense-prediction post-processing hot path of calmiLovesAI/ComputerVision.pytorch (YOLOv8 DFL head
decode, confidence filter, class-aware NMS, and the sibling decoders) as hand-written CUDA kernels
behind a C ABI (include/cvpp.h, libcvpp.so), with a host-side mirror of the reference's Python entry
points under `computervision.pytorch_b200.core`.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from . import distributed  # noqa: F401

__all__ = ["ops", "_lib", "distributed"]
