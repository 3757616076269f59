"""BatchedBlokusEnvironment: drop-in batched counterpart of BlokusEnvironment
(colosseumrl/envs/blokus/BlokusEnvironment.py), backed by csrc/blokus.cuh."""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .base import BatchedBaseEnvironment
from . import _lib

PIECE_NAMES = ["monomino1", "domino1", "trominoe1", "trominoe2", "tetrominoes1", "tetrominoes2", "tetrominoes3",
               "tetrominoes4", "tetrominoes5"] + ["pentominoe%d" % i for i in range(1, 13)]   # board.py:24-44
ORIENTATIONS = ["north", "northeast", "east", "southeast", "south", "southwest", "west", "northwest"]  # board.py:47


def action_to_string(action_id: int) -> str:
    """Engine action id -> the reference's action string (BlokusEnvironment.py:55-80); -1 -> ''."""
    if action_id < 0:
        return ""
    k, o, cell, piece = action_id % 5, (action_id // 5) % 8, (action_id // 40) % 400, action_id // 16000
    return "{};{};{}".format(PIECE_NAMES[piece], (cell % 20, cell // 20), ORIENTATIONS[o] + str(k))


def string_to_action(action_str: str) -> int:
    """The reference's action string (BlokusEnvironment.py:83-106) -> engine action id; '' -> -1."""
    if action_str == "":
        return -1
    piece, index, orientation = action_str.split(";")
    x, y = (int(v) for v in index.replace("(", "").replace(")", "").split(","))
    return ((PIECE_NAMES.index(piece) * 400 + y * 20 + x) * 8 + ORIENTATIONS.index(orientation[:-1])) * 5 + int(orientation[-1])


def _rotate_action_ids(ids, player, inverse: bool):
    """Action ids between the board frame and a player's rotated observation frame (np.rot90(k=-player),
    BlokusEnvironment.py:31-32): index -> M[player] @ (index - 9.5) + 9.5, orientation +- 2 * player, shift unchanged
    (:553-628; matrices board.py:52-73).  Works on ints and on integer tensors / arrays; -1 (pass) stays -1."""
    k, o, cell, piece = ids % 5, (ids // 5) % 8, (ids // 40) % 400, ids // 16000
    x, y = cell % 20, cell // 20
    r = (-player if inverse else player) % 4
    if r == 1:
        x, y = 19 - y, x
    elif r == 2:
        x, y = 19 - x, 19 - y
    elif r == 3:
        x, y = y, 19 - x
    o = (o + (-2 * player if inverse else 2 * player)) % 8
    out = ((piece * 400 + y * 20 + x) * 8 + o) * 5 + k
    if isinstance(ids, int):
        return out if ids >= 0 else -1
    return out * (ids >= 0) - (ids < 0) * 1


def real_action_to_player_perspective(ids, player: int):
    """convert_real_action_to_player_perspective_action (BlokusEnvironment.py:553-588) on action ids."""
    return _rotate_action_ids(ids, int(player), False)


def player_perspective_action_to_real(ids, player: int):
    """convert_player_perspective_action_to_real_action (:591-628) on action ids."""
    return _rotate_action_ids(ids, int(player), True)


@dataclass
class BlokusBatchState:
    packed: torch.Tensor                        # int32 [B, 22, 4]: 352 bytes per game
    result: Optional[torch.Tensor] = None       # uint8 [B, 8] of the step that produced this state


class BatchedBlokusEnvironment(BatchedBaseEnvironment):
    def __init__(self, config: str = "", batch: int = 1, device="cuda:0", seed: int = 0, auto_reset: bool = False,
                 first_env_id: int = 0, capacity: int = 4096):
        super().__init__(config, batch, device, seed, auto_reset, first_env_id)
        self.capacity = int(capacity)      # slots per game in the valid-action list (longest list seen in 16 384 random
                                           # games x 70 steps: 2 104; counts > capacity signals a truncated list)

    @property
    def min_players(self) -> int:
        return 4

    @property
    def max_players(self) -> int:
        return 4

    @property
    def observation_shape(self) -> Dict[str, tuple]:
        return {"board": (20, 20), "pieces": (4, 21), "score": (4,), "player": (1,)}

    @staticmethod
    def observation_names() -> List[str]:
        return ["board", "pieces", "score", "player"]

    @staticmethod
    def all_piece_types() -> List[str]:
        return PIECE_NAMES

    @staticmethod
    def all_orientations() -> List[str]:
        return ORIENTATIONS

    def _alloc(self):
        return torch.empty((self.batch, 22, 4), dtype=torch.int32, device=self.device)

    def new_state(self, num_players: int = 4, out: Optional[BlokusBatchState] = None):
        assert num_players in (None, 4)
        packed = out.packed if out is not None else self._alloc()
        self._check(self._lib.crl_blokus_reset(packed.data_ptr(), None, self.batch, self._stream))
        return BlokusBatchState(packed), torch.ones((self.batch,), dtype=torch.uint8, device=self.device)

    def reset_where(self, state: BlokusBatchState, mask: torch.Tensor) -> BlokusBatchState:
        """new_state for the games with mask != 0 (in place)."""
        mask = self._dev(mask, torch.uint8)
        self._check(self._lib.crl_blokus_reset(state.packed.data_ptr(), mask.data_ptr(), self.batch, self._stream))
        state.result = None
        return state

    def valid_actions(self, state: BlokusBatchState, player: int = -1,
                      out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, count_stats: bool = True):
        """valid_actions (:453-500): (counts int32 [B], ids int32 [B, capacity]) in the reference's order;
        counts == 0 <=> [''].  player = -1: each game's current mover."""
        if out is None:
            out = (torch.empty((self.batch,), dtype=torch.int32, device=self.device),
                   torch.empty((self.batch, self.capacity), dtype=torch.int32, device=self.device))
        counts, ids = out
        self._check(self._lib.crl_blokus_legal(state.packed.data_ptr(), int(player), counts.data_ptr(), ids.data_ptr(),
                                               ids.shape[1], self._stats_ptr if count_stats else None, self.batch,
                                               self.flags, self._stream))
        return counts, ids

    def is_valid_action(self, state: BlokusBatchState, player: int, action) -> torch.Tensor:
        """is_valid_action (:667-719): uint8 [B], 1 iff action[g] is in `player`'s valid list (-1 / '' is not).
        One crl_blokus_is_valid launch: the id is tested against the allowed / anchor boards, no list is built."""
        action = self._dev(action, torch.int32)
        valid = torch.empty((self.batch,), dtype=torch.uint8, device=self.device)
        self._check(self._lib.crl_blokus_is_valid(state.packed.data_ptr(), -1 if player is None else int(player),
                                                  action.data_ptr(), valid.data_ptr(), self.batch, self.flags,
                                                  self._stream))
        return valid

    def player_perspective_valid_actions(self, state: BlokusBatchState, player: int):
        """player_perspective_valid_actions (:502-551): the valid list of `player`, same order, every id rotated into
        that player's observation frame.  (counts, ids) like valid_actions."""
        counts, ids = self.valid_actions(state, player, count_stats=False)
        return counts, real_action_to_player_perspective(ids, player)

    def next_state(self, state: BlokusBatchState, players, actions, out: Optional[BlokusBatchState] = None):
        """next_state (:357-451).  actions: int32 [B] action ids (-1 = pass).
        Returns (new_state, new_players mask, reward int8 [B] (mover's), terminal uint8 [B], winners mask uint8 [B])."""
        new = self.step_(state, actions, out)
        r = new.result
        # views of the record only (bytes 5 / 6 are the players mask and the terminal flag as plain bytes): no launches
        return new, r[:, 5], r[:, 0].view(torch.int8), r[:, 6], r[:, 2]

    def step_(self, state: BlokusBatchState, actions, out: Optional[BlokusBatchState] = None) -> BlokusBatchState:
        """The bare crl_blokus_step launch (out may be `state` itself: in place).  Outputs are in new.result."""
        actions = self._dev(actions, torch.int32)
        new = out if out is not None else BlokusBatchState(self._alloc())
        if new.result is None:
            new.result = self._new_result((self.batch, 8))
        self._check(self._lib.crl_blokus_step(state.packed.data_ptr(), new.packed.data_ptr(), actions.data_ptr(),
                                              new.result.data_ptr(), self._stats_ptr, self.batch, self.flags,
                                              self._stream))
        return new

    def host_stepper(self, state: BlokusBatchState, stream=None) -> "BlokusHostStepper":
        """Two graph launches per game step for a policy that runs on the host (see BlokusHostStepper)."""
        return BlokusHostStepper(self, state, stream=stream)

    def is_terminal(self, state: BlokusBatchState) -> torch.Tensor:
        return ((state.packed[:, 21, 1] >> 16) & 1).to(torch.uint8)

    def current_rewards(self, state: BlokusBatchState) -> torch.Tensor:
        """Scores per player (BlokusEnvironment.py:340-355): int32 [B, 4]."""
        s = state.packed[:, 21, 0]
        return torch.stack([(s >> (8 * p)) & 0xff for p in range(4)], dim=1)

    def state_to_observation(self, state: BlokusBatchState, player: int) -> Dict[str, torch.Tensor]:
        """state_to_observation (:721-768).  player = -1: absolute unpack; -2: every game from its mover's perspective."""
        B = self.batch
        board = torch.empty((B, 20, 20), dtype=torch.int8, device=self.device)
        pieces = torch.empty((B, 4, 21), dtype=torch.uint8, device=self.device)
        score = torch.empty((B, 4), dtype=torch.int32, device=self.device)
        meta = torch.empty((B, 4), dtype=torch.int32, device=self.device)
        self._check(self._lib.crl_blokus_observe(state.packed.data_ptr(), int(player), board.data_ptr(), pieces.data_ptr(),
                                                 score.data_ptr(), meta.data_ptr(), B, self._stream))
        obs = {"board": board, "pieces": pieces, "score": score,
               "player": meta[:, 1:2].clone() if player == -2 else
               torch.full((B, 1), int(player), dtype=torch.int32, device=self.device)}
        if player == -1:
            obs.update(round=meta[:, 0], mover=meta[:, 1], terminal=meta[:, 2], episode_steps=meta[:, 3])
        return obs

    def state_from_arrays(self, board, pieces, score, round_count, mover) -> BlokusBatchState:
        st = BlokusBatchState(self._alloc())
        meta = torch.zeros((self.batch, 4), dtype=torch.int32, device=self.device)
        meta[:, 0] = self._dev(round_count, torch.int32)
        meta[:, 1] = self._dev(mover, torch.int32)
        b, p, s = self._dev(board, torch.int8), self._dev(pieces, torch.uint8), self._dev(score, torch.int32)
        self._check(self._lib.crl_blokus_pack(st.packed.data_ptr(), b.data_ptr(), p.data_ptr(), s.data_ptr(),
                                              meta.data_ptr(), self.batch, self._stream))
        return st

    def random_actions(self, valid, step: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        counts, ids = valid
        out = out if out is not None else torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        self._check(self._lib.crl_blokus_policy_random(counts.data_ptr(), ids.data_ptr(), ids.shape[1], out.data_ptr(),
                                                       self.seed, self.first_env_id, int(step), self.batch, self._stream))
        return out


class BlokusHostStepper:
    """A Blokus batch driven by a policy that runs on the HOST, in two CUDA-graph launches per game step:

      A  `launch_legal()`:  crl_blokus_legal (the movers' ordered valid lists stay on the device) -> D2H of the list
                            LENGTHS into `counts_np` (pinned int32 [B])
      B  `launch_step()`:   H2D of `choice_np` (pinned int32 [B]: index into each game's valid list, negative = pass)
                            -> crl_blokus_pick -> crl_blokus_step (in place on `state`) -> D2H of the result records
                            into `result_np` (pinned uint8 [B, 8])

        stepper = env.host_stepper(state)
        counts = stepper.legal()                     # launch A + wait
        stepper.choice_np[...] = my_policy(counts)   # e.g. a random index < counts
        records = stepper.step()                     # launch B + wait
    The kilobytes of action ids per game never cross PCIe; an actor serving several batches pipelines them
    (`launch_legal` / `wait_legal` / `launch_step` / `wait_step`), one stream per stepper.  `ids_device` is the
    device-side list for policies that want to look at the ids themselves."""

    def __init__(self, env, state, stream=None):
        from . import _cudart
        self.env, self.state = env, state
        B, dev = env.batch, env.device
        with torch.cuda.device(dev):
            s = stream if stream is not None else torch.cuda.current_stream(dev)
            self.counts = torch.zeros((B,), dtype=torch.int32).pin_memory()
            self.choice = torch.zeros((B,), dtype=torch.int32).pin_memory()
            self.result = torch.zeros((B, 8), dtype=torch.uint8).pin_memory()
            self.counts_np, self.choice_np, self.result_np = self.counts.numpy(), self.choice.numpy(), self.result.numpy()
            self._counts_dev = torch.zeros((B,), dtype=torch.int32, device=dev)
            self.ids_device = torch.zeros((B, env.capacity), dtype=torch.int32, device=dev)
            self._choice_dev = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._action_dev = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._result_dev = torch.zeros((B, 8), dtype=torch.uint8, device=dev)

            def phase_a():
                env.valid_actions(state, -1, out=(self._counts_dev, self.ids_device))
                self.counts.copy_(self._counts_dev, non_blocking=True)

            def phase_b():
                self._choice_dev.copy_(self.choice, non_blocking=True)
                env._check(env._lib.crl_blokus_pick(self._counts_dev.data_ptr(), self.ids_device.data_ptr(), env.capacity,
                                                    self._choice_dev.data_ptr(), self._action_dev.data_ptr(), B, env._stream))
                env._check(env._lib.crl_blokus_step(state.packed.data_ptr(), state.packed.data_ptr(), self._action_dev.data_ptr(),
                                                    self._result_dev.data_ptr(), env._stats_ptr, B, env.flags, env._stream))
                state.result = None
                self.result.copy_(self._result_dev, non_blocking=True)

            with torch.cuda.stream(s):
                phase_a()                                        # warm the launch paths (the state is not stepped:
            torch.cuda.synchronize(dev)                          # phase B is only captured, never run eagerly)
            self._graphs = []
            for fn in (phase_a, phase_b):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                self._graphs.append(g)
            raw = _lib.load()
            self._launch_fn, self._wait_fn = raw.crl_host_graph_launch, raw.crl_host_event_wait
            self._exec = [g.raw_cuda_graph_exec() for g in self._graphs]
            self._done = [_cudart.new_event(), _cudart.new_event()]
            self._stream_handle = s.cuda_stream
            for e in self._exec:
                _cudart.check(_cudart.rt().cudaGraphUpload(e, self._stream_handle), "cudaGraphUpload")

    def _launch(self, i):
        rc = self._launch_fn(self._exec[i], self._stream_handle, self._done[i])
        if rc == _lib.ERR_ARG:                                   # wrong current device: retry under a guard
            with torch.cuda.device(self.env.device):
                rc = self._launch_fn(self._exec[i], self._stream_handle, self._done[i])
        if rc:
            raise RuntimeError("BlokusHostStepper launch failed: %s" % _lib.load().crl_last_error().decode())

    def _wait(self, i):
        if self._wait_fn(self._done[i]):
            raise RuntimeError("BlokusHostStepper wait failed: %s" % _lib.load().crl_last_error().decode())

    def launch_legal(self):
        self._launch(0)

    def wait_legal(self):
        self._wait(0)
        return self.counts_np

    def launch_step(self):
        self._launch(1)

    def wait_step(self):
        self._wait(1)
        return self.result_np

    def legal(self):
        self.launch_legal()
        return self.wait_legal()

    def step(self):
        self.launch_step()
        return self.wait_step()
