"""sow_b200 -- B200-native (sm_100a) implementation of the SoW ("Sum-of-Weights") training hot path.

The public surface is the reference's own: import ``tn_gradient`` (a thin re-export package in this repo) exactly
as with antoine311200/sow.  All device math lives in ``sow_b200/csrc`` (hand-written CUDA behind the C ABI declared
in ``include/sow_b200.h``); there is no CPU fallback -- the ops raise when the extension or a CUDA device is missing.
"""
from ._lib import SowB200Error, LIB_PATH  # noqa: F401

__all__ = ["SowB200Error", "LIB_PATH"]
__version__ = "0.1.0"
