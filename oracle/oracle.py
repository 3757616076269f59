"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg may
import this module.  The product package ``colosseumrl_b200`` never does.

Every function takes / returns plain numpy arrays in the *reference's* unpacked layout (see the
headers of oracle_tron.c / oracle_ttt.c / oracle_blokus.c) so results can be compared directly with
states extracted from the reference's own Python objects (tests/golden/*.npz).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NSTAT = 32

PIECE_NAMES = ["monomino1", "domino1", "trominoe1", "trominoe2", "tetrominoes1", "tetrominoes2",
               "tetrominoes3", "tetrominoes4", "tetrominoes5"] + ["pentominoe%d" % i for i in range(1, 13)]
ORIENTATIONS = ["north", "northeast", "east", "southeast", "south", "southwest", "west", "northwest"]


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc) if missing or stale; returns its path."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_tron.c", "oracle_ttt.c", "oracle_blokus.c", "oracle_rollout.c")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        cc = os.environ.get("CC", "gcc")
        cmd = [cc, "-O2", "-fPIC", "-pthread", "-shared", "-o", so] + srcs + ["-lm"]
        subprocess.check_call(cmd)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i64(a):
    return _p(a, C.c_int64)


def num_threads() -> int:
    return lib().orc_num_threads()


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


# ----------------------------------------------------------------------------- Tron
def tron_start_positions(N, P, ring_offset=1, spawn_offset=2):
    heads = np.zeros(P, np.int64)
    dirs = np.zeros(P, np.int64)
    rc = lib().orc_tron_start_positions(N, P, ring_offset, spawn_offset, _i64(heads), _i64(dirs))
    if rc:
        raise ValueError("bad tron config")
    return heads, dirs


def tron_new_state(N, P):
    board = np.zeros((N, N), np.int64)
    heads = np.zeros(P, np.int64)
    dirs = np.zeros(P, np.int64)
    deaths = np.zeros(P, np.int64)
    rc = lib().orc_tron_new_state(N, P, _i64(board), _i64(heads), _i64(dirs), _i64(deaths))
    if rc:
        raise ValueError("bad tron config")
    return board, heads, dirs, deaths


def tron_next_state(state, actions):
    """Functional: returns (new_state, alive_mask, rewards[P], terminal, winners_mask)."""
    board, heads, dirs, deaths = (np.array(a, np.int64, copy=True) for a in state)
    N, P = board.shape[0], heads.shape[0]
    act = np.ascontiguousarray(actions, np.int64)
    rewards = np.zeros(P, np.int64)
    alive = C.c_int(0)
    winners = C.c_int(0)
    term = lib().orc_tron_next_state(N, P, _i64(board), _i64(heads), _i64(dirs), _i64(deaths), _i64(act),
                                     _i64(rewards), C.byref(alive), C.byref(winners))
    return (board, heads, dirs, deaths), alive.value, rewards, bool(term), winners.value


def tron_compute_ranking(state):
    board, _, _, deaths = state
    board = np.ascontiguousarray(board, np.int64)
    deaths = np.ascontiguousarray(deaths, np.int64)
    rank = np.zeros(deaths.shape[0], np.int64)
    lib().orc_tron_compute_ranking(board.shape[0], deaths.shape[0], _i64(board), _i64(deaths), _i64(rank))
    return rank


def tron_observation(state, player):
    board, heads, dirs, deaths = (np.ascontiguousarray(a, np.int64) for a in state)
    N, P = board.shape[0], heads.shape[0]
    ob, oh, od, ode = np.zeros_like(board), np.zeros_like(heads), np.zeros_like(dirs), np.zeros_like(deaths)
    lib().orc_tron_observation(N, P, player, _i64(board), _i64(heads), _i64(dirs), _i64(deaths),
                               _i64(ob), _i64(oh), _i64(od), _i64(ode))
    return {"board": ob, "heads": oh, "directions": od, "deaths": ode}


class TronBatch:
    """Unpacked batched Tron state for rollouts (arrays [B, ...])."""

    def __init__(self, B, N=19, P=4):
        self.B, self.N, self.P = B, N, P
        self.board = np.zeros((B, N, N), np.int64)
        self.heads = np.zeros((B, P), np.int64)
        self.directions = np.zeros((B, P), np.int64)
        self.deaths = np.zeros((B, P), np.int64)
        self.terminal = np.zeros(B, np.uint8)
        self.ep_len = np.zeros(B, np.int32)
        self.stats = np.zeros(NSTAT, np.int64)

    def rollout(self, seed, env0, t0, K, fresh=False, nthreads=0):
        lib().orc_tron_rollout(self.N, self.P, C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.B),
                               C.c_uint32(t0), K, int(fresh), _i64(self.board), _i64(self.heads),
                               _i64(self.directions), _i64(self.deaths), _p(self.terminal, C.c_uint8),
                               _p(self.ep_len, C.c_int32), _i64(self.stats), nthreads)


# ----------------------------------------------------------------------------- TTT
TTT_SHAPE = {2: (3, 3), 3: (3, 5), 4: (3, 3, 3)}


def ttt_cells(n):
    return lib().orc_ttt_cells(n)


def ttt_lines(n):
    out = np.zeros((64, 3), np.int32)
    cnt = lib().orc_ttt_lines(n, _p(out, C.c_int32))
    return out[:cnt].copy()


def ttt_new_state(n):
    return np.full(TTT_SHAPE[n], -1, np.int8), -1


def ttt_next_state(n, state, player, action):
    """action: flat C-order cell index or -1 ('').  Returns (new_state, next_player, reward, terminal, winner|-1)."""
    board, winner = state
    board = np.array(board, np.int8, copy=True)
    w = C.c_int(int(winner))
    reward, term, wout = C.c_int(0), C.c_int(0), C.c_int(0)
    nxt = lib().orc_ttt_next_state(n, _p(board, C.c_int8), C.byref(w), int(player), int(action),
                                   C.byref(reward), C.byref(term), C.byref(wout))
    return (board, w.value), nxt, reward.value, bool(term.value), wout.value


def ttt_valid_actions(n, state):
    board = np.ascontiguousarray(state[0], np.int8)
    out = np.zeros(32, np.int32)
    k = lib().orc_ttt_valid_actions(n, _p(board, C.c_int8), _p(out, C.c_int32))
    return out[:k].copy()


def ttt_observation(n, state, player):
    board = np.ascontiguousarray(state[0], np.int8)
    out = np.zeros_like(board)
    lib().orc_ttt_observation(n, _p(board, C.c_int8), int(player), _p(out, C.c_int8))
    return out


class TTTBatch:
    def __init__(self, B, n):
        self.B, self.n = B, n
        self.board = np.full((B,) + TTT_SHAPE[n], -1, np.int8)
        self.winner = np.full(B, -1, np.int32)
        self.mover = np.zeros(B, np.int32)
        self.terminal = np.zeros(B, np.uint8)
        self.ep_len = np.zeros(B, np.int32)
        self.stats = np.zeros(NSTAT, np.int64)

    def rollout(self, seed, env0, t0, K, fresh=False, nthreads=0):
        lib().orc_ttt_rollout(self.n, C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.B), C.c_uint32(t0), K,
                              int(fresh), _p(self.board, C.c_int8), _p(self.winner, C.c_int32),
                              _p(self.mover, C.c_int32), _p(self.terminal, C.c_uint8), _p(self.ep_len, C.c_int32),
                              _i64(self.stats), nthreads)


# ----------------------------------------------------------------------------- Blokus
def blokus_encode_action(piece, x, y, orient, k):
    return ((piece * 400 + y * 20 + x) * 8 + orient) * 5 + k


def blokus_decode_action(aid):
    aid = int(aid)
    return aid // 16000, (aid // 40) % 20, (aid // 40) % 400 // 20, (aid // 5) % 8, aid % 5  # piece, x, y, o, k


def blokus_action_to_string(aid):
    """Reference string form '{piece};{(x, y)};{orient}{k}' (BlokusEnvironment.py:55-80); '' for pass."""
    if aid < 0:
        return ""
    piece, x, y, o, k = blokus_decode_action(aid)
    return "{};{};{}".format(PIECE_NAMES[piece], (x, y), ORIENTATIONS[o] + str(k))


def blokus_string_to_action(s):
    if s == "":
        return -1
    piece, index, orient = s.split(";")
    x, y = (int(v) for v in index.replace("(", "").replace(")", "").split(","))
    return blokus_encode_action(PIECE_NAMES.index(piece), x, y, ORIENTATIONS.index(orient[:-1]), int(orient[-1]))


def blokus_new_state():
    return (np.zeros((20, 20), np.int64), 0, np.ones((4, 21), np.uint8), np.zeros(4, np.int64))


def blokus_anchors(board, round_count, color):
    board = np.ascontiguousarray(board, np.int64)
    out = np.zeros((400, 2), np.int32)
    n = lib().orc_blokus_anchors(_i64(board), int(round_count), int(color), _p(out, C.c_int32))
    return out[:n].copy()


def blokus_valid_moves(state, player, cap=16384):
    board, round_count, inv, _ = state
    board = np.ascontiguousarray(board, np.int64)
    invp = np.ascontiguousarray(inv[player], np.uint8)
    out = np.zeros(cap, np.int32)
    n = lib().orc_blokus_valid_moves(_i64(board), int(round_count), int(player) + 1, _p(invp, C.c_uint8),
                                     _p(out, C.c_int32), cap)
    assert n <= cap
    return out[:n].copy()


def blokus_next_state(state, mover, action_id):
    """Functional: returns (new_state, next_mover, reward, terminal, winners_mask)."""
    board, round_count, inv, scores = state
    board = np.array(board, np.int64, copy=True)
    inv = np.array(inv, np.uint8, copy=True)
    scores = np.array(scores, np.int64, copy=True)
    rc = C.c_int(int(round_count))
    reward, term, winners = C.c_int(0), C.c_int(0), C.c_int(0)
    nxt = lib().orc_blokus_next_state(_i64(board), C.byref(rc), _p(inv, C.c_uint8), _i64(scores), int(mover),
                                      int(action_id), C.byref(reward), C.byref(term), C.byref(winners))
    return (board, rc.value, inv, scores), nxt, reward.value, bool(term.value), winners.value


def blokus_observation(state, player):
    board, _, inv, scores = state
    board = np.ascontiguousarray(board, np.int64)
    inv = np.ascontiguousarray(inv, np.uint8)
    scores = np.ascontiguousarray(scores, np.int64)
    ob = np.zeros((20, 20), np.int64)
    op = np.zeros((4, 21), np.uint8)
    osc = np.zeros(4, np.int64)
    lib().orc_blokus_observation(_i64(board), _p(inv, C.c_uint8), _i64(scores), int(player), _i64(ob),
                                 _p(op, C.c_uint8), _i64(osc))
    return {"board": ob, "pieces": op, "score": osc, "player": np.array([player])}


class BlokusBatch:
    def __init__(self, B):
        self.B = B
        self.board = np.zeros((B, 20, 20), np.int64)
        self.round_count = np.zeros(B, np.int32)
        self.inventory = np.ones((B, 4, 21), np.uint8)
        self.scores = np.zeros((B, 4), np.int64)
        self.mover = np.zeros(B, np.int32)
        self.terminal = np.zeros(B, np.uint8)
        self.ep_len = np.zeros(B, np.int32)
        self.stats = np.zeros(NSTAT, np.int64)

    def rollout(self, seed, env0, t0, K, fresh=False, nthreads=0):
        lib().orc_blokus_rollout(C.c_uint64(seed), C.c_int64(env0), C.c_int64(self.B), C.c_uint32(t0), K,
                                 int(fresh), _i64(self.board), _p(self.round_count, C.c_int32),
                                 _p(self.inventory, C.c_uint8), _i64(self.scores), _p(self.mover, C.c_int32),
                                 _p(self.terminal, C.c_uint8), _p(self.ep_len, C.c_int32), _i64(self.stats),
                                 nthreads)
