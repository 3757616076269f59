"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* reference (container only).

Imports the unmodified reference game classes from ``/root/reference`` without
importing ``colosseumrl/__init__.py`` (which pulls in spacetime / gym / ray /
pygame, none of which exist here).  Only ``oracle/make_golden.py`` and the
container-only cross-check tests use this; nothing on the GPU box does
(``/root/reference`` is absent there).

Recipe:
  1. ``oracle/_ref/CyTronGrid*.so`` is the reference's own Cython source
     (envs/tron/CyTronGrid.pyx) compiled where it lies by ``oracle/Makefile``.
  2. Stub *packages* (``types.ModuleType`` with ``__path__``) are registered for
     ``colosseumrl`` and its sub-packages so that leaf modules import directly.
  3. A stub ``pygame`` satisfies ``envs/blokus/gui.py:13-16``.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("COLOSSEUM_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_BUILD = os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "colosseumrl", "envs"))


def _stub_pkg(name, paths):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = list(paths)
    m.__package__ = name
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return dict of reference classes/modules. Raises if the reference is absent."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    pkg = os.path.join(REF_ROOT, "colosseumrl")
    _stub_pkg("colosseumrl", [pkg])
    _stub_pkg("colosseumrl.envs", [os.path.join(pkg, "envs")])
    _stub_pkg("colosseumrl.envs.tron", [os.path.join(pkg, "envs", "tron"), _REF_BUILD])
    _stub_pkg("colosseumrl.envs.blokus", [os.path.join(pkg, "envs", "blokus")])
    _stub_pkg("colosseumrl.envs.tictactoe", [os.path.join(pkg, "envs", "tictactoe")])
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        pg.time = types.SimpleNamespace(Clock=lambda: None)
        sys.modules["pygame"] = pg

    import importlib
    base = importlib.import_module("colosseumrl.BaseEnvironment")
    tron = importlib.import_module("colosseumrl.envs.tron.TronGridEnvironment")
    blokus = importlib.import_module("colosseumrl.envs.blokus.BlokusEnvironment")
    board = importlib.import_module("colosseumrl.envs.blokus.board")
    ai = importlib.import_module("colosseumrl.envs.blokus.ai")
    t2 = importlib.import_module("colosseumrl.envs.tictactoe.tictactoe_2p_env")
    t3 = importlib.import_module("colosseumrl.envs.tictactoe.tictactoe_3p_env")
    t4 = importlib.import_module("colosseumrl.envs.tictactoe.tictactoe_4p_env")
    _loaded.update(
        BaseEnvironment=base.BaseEnvironment,
        TronGridEnvironment=tron.TronGridEnvironment,
        BlokusEnvironment=blokus.BlokusEnvironment,
        blokus_module=blokus,
        blokus_board=board,
        blokus_ai=ai,
        TicTacToe2PlayerEnv=t2.TicTacToe2PlayerEnv,
        TicTacToe3PlayerEnv=t3.TicTacToe3PlayerEnv,
        TicTacToe4PlayerEnv=t4.TicTacToe4PlayerEnv,
        ttt_modules={2: t2, 3: t3, 4: t4},
    )
    return _loaded
