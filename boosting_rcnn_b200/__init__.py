"""boosting_rcnn_b200 — B200-native (sm_100a) proposal-to-RoI hot path of
Boosting R-CNN behind the mmdet registry names (ATSSRPNHead,
SingleRoIExtractor, ProbRoIHead, ProbConvFCBBoxHead).

Compute lives in ``libbrcnn.so`` (hand-written CUDA, C ABI in
``include/brcnn.h``); this package is the host-side mirror of the reference
interfaces.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'


def build(verbose=False):
    """Compile libbrcnn.so in-tree (nvcc, sm_100a)."""
    return _lib.build(verbose=verbose)
