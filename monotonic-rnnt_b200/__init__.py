"""monotonic-rnnt_b200: B200-native (sm_100a) loss-and-gradient path of the monotonic RNN-T loss.

The package is a thin host-side mirror of the reference's PyTorch operator interface
(pytorch_binding/monotonic_rnnt_op.py) over the flat C ABI of libmonotonic_rnnt.so
(include/mrnnt_c_api.h).  The directory name carries a hyphen, so import it as
``monotonic_rnnt_b200`` (the shim module of that name at the repository root does the loading).
"""
from . import _lib, build, peer, shard, synth  # noqa: F401
from ._lib import RNNTError  # noqa: F401
from .rnnt_op import (  # noqa: F401
    LossHandle,
    MonotonicRNNTFunction,
    MonotonicRNNTLoss,
    monotonic_rnnt_loss,
    workspace_size,
)

__all__ = ["LossHandle", "MonotonicRNNTFunction", "MonotonicRNNTLoss", "monotonic_rnnt_loss", "workspace_size",
           "RNNTError", "build", "peer", "shard", "synth"]
