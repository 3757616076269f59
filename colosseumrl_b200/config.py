"""Name -> batched environment class registry, same names as colosseumrl/config.py:37-44."""
from typing import Callable, Dict, List


def _tron():
    from .tron import BatchedTronGridEnvironment
    return BatchedTronGridEnvironment


def _blokus():
    from .blokus import BatchedBlokusEnvironment
    return BatchedBlokusEnvironment


def _ttt(n):
    def f():
        from . import tictactoe
        return getattr(tictactoe, "BatchedTicTacToe%dPlayerEnv" % n)
    return f


ENVIRONMENT_CLASSES: Dict[str, Callable] = {
    "blokus": _blokus,
    "tron": _tron,
    "tictactoe": _ttt(2),
    "tictactoe_3p": _ttt(3),
    "tictactoe_4p": _ttt(4),
}


def get_environment(environment: str):
    return ENVIRONMENT_CLASSES[environment]()


def available_environments() -> List[str]:
    return list(ENVIRONMENT_CLASSES.keys())
