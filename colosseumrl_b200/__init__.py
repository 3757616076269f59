"""colosseumrl_b200 -- batched, bit-packed, B200-native (sm_100a) game dynamics for ColosseumRL.

Importing this package does not need a GPU; constructing an environment does (there is no CPU fallback).
"""
from ._lib import CrlError, FLAG_AUTO_RESET, NSTAT  # noqa: F401


def __getattr__(name):
    # lazy: the environment classes import torch
    if name in ("BatchedBaseEnvironment",):
        from .base import BatchedBaseEnvironment
        return BatchedBaseEnvironment
    if name in ("BatchedTronGridEnvironment", "TronBatchState"):
        from . import tron
        return getattr(tron, name)
    raise AttributeError(name)
