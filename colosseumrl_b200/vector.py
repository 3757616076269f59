"""Vectorised reset / step wrapper: the batched analogue of the reference's RLlib adapter
(colosseumrl/envs/wrappers/rllib.py:8-55, envs/tron/rllib.py:13-66), which drives ONE game per Python object through
new_state -> state_to_observation and next_state -> state_to_observation -> reward / done dicts.  Here one object
steps B games per call and the per-agent dicts become tensors:

    venv = VectorEnv(BatchedTronGridEnvironment("", batch=4096, auto_reset=True))
    obs = venv.reset()                       # {"board": [B, P, N, N], "heads": [B, P, P], ...}
    obs, rewards, dones, info = venv.step(actions)     # rewards [B, P], dones uint8 [B]

Simultaneous-move games (Tron) are observed from every seat; turn-based games (Blokus, Tic Tac Toe) from the seat of
each game's next mover (`info["mover"]`).  `dones` marks the step that ended an episode.  With auto_reset=True a
finished game is restarted INSIDE that step() call (gym's vector-env convention): the observation, `info["mover"]` and
`valid_actions()` returned next all describe the fresh game the next action will be applied to, never the finished
board; `step(..., final_observation=True)` additionally returns the finished games' last observation in
`info["final_observation"]`.
"""
from typing import Dict

import torch


class VectorEnv:
    def __init__(self, env):
        self.env = env
        self.state = None
        self.simultaneous = hasattr(env, "remove_on_death")          # Tron: all live players act every step

    @property
    def num_envs(self) -> int:
        return self.env.batch

    @property
    def num_players(self) -> int:
        return self.env.max_players

    def _observe(self) -> Dict[str, torch.Tensor]:
        if self.simultaneous:
            return self.env.state_to_observation(self.state, -3)      # CRL_PLAYER_ALL: every seat in one launch
        return self.env.state_to_observation(self.state, -2)          # CRL_PLAYER_MOVER

    def reset(self) -> Dict[str, torch.Tensor]:
        self.state, _ = self.env.new_state()
        return self._observe()

    def valid_actions(self):
        """Valid actions of every game's acting player(s), in the batched environment's own format."""
        return self.env.valid_actions(self.state, -1)

    def step(self, actions, final_observation: bool = False):
        """actions: the batched environment's action tensor (Tron int8 [B, 4]; Blokus int32 [B] ids; TTT int8 [B])."""
        new, players, rewards, terminal, winners = self.env.next_state(self.state, None, actions, out=self.state)
        self.state = new
        info = {"players": players, "winners": winners}
        if self.env.auto_reset:
            if final_observation:
                info["final_observation"] = self._observe()
            # restart the finished games now, so that what the agent sees next is what its next action applies to
            self.env.reset_where(self.state, terminal)
            fresh = terminal != 0
            first = (1 << self.num_players) - 1 if self.simultaneous else 1          # new_state's acting players
            players = torch.where(fresh, torch.full_like(players, first), players)
            info["players"] = players
        if not self.simultaneous:                                     # players = 1 << next mover
            info["mover"] = (players >> 1) - (players >> 3)           # 1, 2, 4, 8 -> 0, 1, 2, 3
        return self._observe(), rewards, terminal, info
