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
    if name in ("BatchedTicTacToe2PlayerEnv", "BatchedTicTacToe3PlayerEnv", "BatchedTicTacToe4PlayerEnv", "TTTBatchState"):
        from . import tictactoe
        return getattr(tictactoe, name)
    if name in ("BatchedBlokusEnvironment", "BlokusBatchState"):
        from . import blokus
        return getattr(blokus, name)
    if name == "get_environment":
        from .config import get_environment
        return get_environment
    raise AttributeError(name)
