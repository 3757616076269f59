"""TEST INFRASTRUCTURE ONLY -- widen the reference-recorded Blokus pin (tests/golden/blokus_wide.npz).

Run in the build container (where /root/reference exists):

    python oracle/make_golden_wide.py [n_games]

The unmodified reference ``BlokusEnvironment`` (through oracle/ref_shim.py) is driven exactly as in
oracle/make_golden.py, but recorded compactly so that many more games fit a small fixture:

* ``games``: 64 full random games (Philox seed 1).  Per step: mover, chosen action id, the LENGTH and an
  order-sensitive 64-bit hash of the reference's ordered ``valid_actions`` list (``list_hash`` below; the first 8 games
  of tests/golden/blokus_games.npz keep the lists themselves), and every ``next_state`` output with the full state.
* ``end``: hand-built end-game positions the random games hardly ever reach -- taken from mid / late positions of
  those games with the inventories cut down to 0..2 pieces (the monomino as the LAST piece: +20, any other last piece:
  +15, ``ai.py:49-54``), scores forced to ties (``BlokusEnvironment.py:426-437``: several winners; the reward is the
  mover's index in the stable sort by score, ``:439-440``) -- input state, valid list (length + hash), action, outputs.

Compared with the C oracle in tests/test_oracle_golden.py and with the CUDA engine in tests/cases_blokus.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import PIECE_NAMES, blokus_action_to_string, blokus_string_to_action  # noqa: E402
from colosseumrl_b200.philox import env_step_words, TAG_BLOKUS  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
HASH_M = np.uint64(0x9E3779B97F4A7C15)


def list_hash(ids):
    """Order-sensitive hash of an id list: sum_i (id_i + 1) * M^(i+1) mod 2^64."""
    ids = np.asarray(ids, np.int64)
    if ids.size == 0:
        return np.uint64(0)
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.full(ids.size, HASH_M, np.uint64))
        return ((ids + 1).astype(np.uint64) * pw).sum(dtype=np.uint64)


def state_arrays(state):
    board, round_count, players = state
    inv = np.zeros((4, 21), np.uint8)
    for i, p in enumerate(players):
        for name in p.current_pieces:
            inv[i, PIECE_NAMES.index(name)] = 1
    scores = np.asarray([p.player_score for p in players], np.int64)
    return board.board_contents.astype(np.int8), int(round_count), inv, scores


def valid_ids(env, state, mover):
    va = env.valid_actions(state, mover)
    ids = [blokus_string_to_action(s) for s in va if s != ""]
    assert all(blokus_action_to_string(i) == s for i, s in zip(ids, va))
    return ids


def winners_mask(winners):
    return sum(1 << int(p) for p in winners) if winners is not None else 0


def play_games(R, seed, n_games, keep_states):
    env = R["BlokusEnvironment"]()
    rec = {k: [] for k in ("game", "t", "mover", "action", "n_valid", "valid_hash", "board", "round", "inventory",
                           "scores", "reward", "terminal", "winners", "next_mover")}
    for g in range(n_games):
        state, players = env.new_state()
        after = 0
        for t in range(200):
            mover = players[0]
            r = env_step_words(seed, [1000 + g], t, TAG_BLOKUS)[0]
            ids = valid_ids(env, state, mover)
            action = ids[int(r[0] % len(ids))] if ids else -1
            if t >= 24 and (t + g) % 5 == 0:
                keep_states.append((state, mover))
            state, players, rewards, terminal, winners = env.next_state(state, [mover], [blokus_action_to_string(action)])
            b, rc, inv, sc = state_arrays(state)
            rec["game"].append(g); rec["t"].append(t); rec["mover"].append(mover); rec["action"].append(action)
            rec["n_valid"].append(len(ids)); rec["valid_hash"].append(list_hash(ids))
            rec["board"].append(b); rec["round"].append(rc); rec["inventory"].append(inv); rec["scores"].append(sc)
            rec["reward"].append(rewards[0]); rec["terminal"].append(bool(terminal))
            rec["winners"].append(winners_mask(winners)); rec["next_mover"].append(players[0])
            if terminal:
                after += 1
                if after > 1:
                    break
        print("  game", g, "steps", t + 1, "scores", sc, flush=True)
    return {k: np.asarray(v) for k, v in rec.items()}


def end_positions(R, keep_states, seed=11):
    """Inventories cut down / scores tied on reference-reached boards, then ONE reference step each."""
    from copy import deepcopy
    env = R["BlokusEnvironment"]()
    values = R["blokus_ai"].GAME_PIECE_VALUES
    rng = np.random.RandomState(seed)
    rec = {k: [] for k in ("i_board", "i_round", "i_inventory", "i_scores", "mover", "action", "n_valid", "valid_hash",
                           "board", "round", "inventory", "scores", "reward", "terminal", "winners", "next_mover")}
    for n, (state, mover) in enumerate(keep_states):
        board, round_count, players = state
        players = deepcopy(players)
        kind = n % 4
        for q, p in enumerate(players):
            have = list(p.current_pieces)
            if q == mover:
                if kind == 0:
                    keep = ["monomino1"]                               # last piece = monomino: +20
                elif kind == 1:
                    keep = [have[rng.randint(len(have))]] if have else []   # any last piece: +15 (or +20)
                else:
                    keep = [have[i] for i in sorted(rng.choice(len(have), size=min(2, len(have)), replace=False))]
                if kind == 3:
                    keep = keep[:1]
            else:
                k = 0 if kind == 3 else int(rng.randint(0, 3))           # kind 3: nobody else holds a piece -> the game ends
                keep = [have[i] for i in sorted(rng.choice(len(have), size=min(k, len(have)), replace=False))] if have and k else []
            p.current_pieces = [x for x in PIECE_NAMES if x in keep]   # inventory order = the reference's
            if kind == 3:                                              # ties: several winners (incl. the mover after its bonus)
                p.player_score = int(rng.choice([44, 45]))
                if q == mover and keep and rng.rand() < 0.6:
                    gain = values[keep[0]] + (20 if keep[0] == "monomino1" else 15)
                    p.player_score = 45 - gain
            elif kind == 2:
                p.player_score = int(rng.randint(0, 90))
        st = (board, round_count, players)
        ids = valid_ids(env, st, mover)
        action = ids[int(rng.randint(len(ids)))] if ids and rng.rand() < 0.9 else -1
        ib, irc, iinv, isc = state_arrays(st)
        nst, nplayers, rewards, terminal, winners = env.next_state(st, [mover], [blokus_action_to_string(action)])
        b, rc, inv, sc = state_arrays(nst)
        rec["i_board"].append(ib); rec["i_round"].append(irc); rec["i_inventory"].append(iinv); rec["i_scores"].append(isc)
        rec["mover"].append(mover); rec["action"].append(action); rec["n_valid"].append(len(ids))
        rec["valid_hash"].append(list_hash(ids)); rec["board"].append(b); rec["round"].append(rc)
        rec["inventory"].append(inv); rec["scores"].append(sc); rec["reward"].append(rewards[0])
        rec["terminal"].append(bool(terminal)); rec["winners"].append(winners_mask(winners)); rec["next_mover"].append(nplayers[0])
    return {k: np.asarray(v) for k, v in rec.items()}


def main():
    n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    R = ref_shim.load()
    keep = []
    g = play_games(R, seed=1, n_games=n_games, keep_states=keep)
    e = end_positions(R, keep)
    out = {"g_" + k: v for k, v in g.items()}
    out.update({"e_" + k: v for k, v in e.items()})
    np.savez_compressed(os.path.join(OUT, "blokus_wide.npz"), **out)
    bonus = (e["scores"].sum(1) - e["i_scores"].sum(1))
    print("games", n_games, "steps", len(g["t"]), "terminal", int(g["terminal"].sum()), "max valid", int(g["n_valid"].max()))
    print("end positions", len(e["mover"]), "terminal", int(e["terminal"].sum()),
          "multi-winner", int(sum(bin(int(w)).count("1") > 1 for w in e["winners"])),
          "bonus steps (>5 points)", int((bonus > 5).sum()))


if __name__ == "__main__":
    main()
