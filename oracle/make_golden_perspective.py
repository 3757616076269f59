"""TEST INFRASTRUCTURE ONLY -- tests/golden/blokus_perspective.npz from the REAL reference: the player-perspective
action transforms of BlokusEnvironment (convert_real_action_to_player_perspective_action :553-588,
convert_player_perspective_action_to_real_action :591-628, player_perspective_valid_actions :502-551).

    python oracle/make_golden_perspective.py        (build container only: needs /root/reference)
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.oracle import blokus_action_to_string, blokus_string_to_action  # noqa: E402


def to_id(s):
    """Action string -> engine id.  Under numpy >= 2 the reference formats the rotated index as
    '(np.int32(6), np.int32(4))' (same family as SURVEY.md Appendix A-X1): normalise before parsing."""
    if s == "":
        return -1
    piece, index, orientation = s.split(";")
    nums = re.findall(r"\((-?\d+)\)", index) if "np." in index else re.findall(r"-?\d+", index)
    return blokus_string_to_action("%s;(%d, %d);%s" % (piece, int(nums[0]), int(nums[1]), orientation))


def main():
    R = ref_shim.load()
    env = R["BlokusEnvironment"]()
    rng = np.random.RandomState(3)
    n = 4000
    piece, x, y = rng.randint(0, 21, n), rng.randint(0, 20, n), rng.randint(0, 20, n)
    o, player = rng.randint(0, 8, n), rng.randint(0, 4, n)
    sizes = [1, 2, 3, 3, 4, 4, 4, 4, 4] + [5] * 12
    k = np.array([rng.randint(0, sizes[p]) for p in piece])
    real = ((piece * 400 + y * 20 + x) * 8 + o) * 5 + k
    persp, back = [], []
    for a, pl in zip(real, player):
        s = env.convert_real_action_to_player_perspective_action(blokus_action_to_string(int(a)), int(pl))
        persp.append(to_id(s))
        clean = blokus_action_to_string(persp[-1])
        back.append(to_id(env.convert_player_perspective_action_to_real_action(clean, int(pl))))
    # player_perspective_valid_actions on a few recorded positions
    g = np.load(os.path.join(ROOT, "tests", "golden", "blokus_games.npz"))
    Board, AI = R["blokus_board"].Board, R["blokus_ai"].AI
    names = env.all_piece_types()
    idx = [0, 1, 2, 3, 9, 22, 35, 48]
    flat, off, who = [], [0], []
    for i in idx:
        board = Board()
        board.board_contents = g["board"][i].astype(np.int64)
        players = []
        for q in range(4):
            ai = AI(board, q + 1)
            ai.current_pieces = [nm for j, nm in enumerate(names) if g["inventory"][i][q, j]]
            ai.player_score = int(g["scores"][i][q])
            players.append(ai)
        state = (board, int(g["round"][i]), players)
        p = int(g["next_mover"][i])
        lst = env.player_perspective_valid_actions(state, p)
        flat += [to_id(s) for s in lst]
        off.append(len(flat))
        who.append(p)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "blokus_perspective.npz"), real=real, player=player,
                        perspective=np.array(persp), back=np.array(back), pos_idx=np.array(idx), pos_player=np.array(who),
                        pos_flat=np.array(flat, np.int64), pos_off=np.array(off))
    print("wrote blokus_perspective.npz", n, len(flat))


if __name__ == "__main__":
    main()
