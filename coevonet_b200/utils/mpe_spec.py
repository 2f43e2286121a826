"""Host-side description of ``simple_adversary_v3`` for the device rollout:
agent order, spaces, and the reference-compatible initial-state stream.

The physics itself runs inside the K1 kernels (``csrc/common.cuh``); this file
only reproduces what the reference draws on the HOST per ``env.reset()``
(upstream ``reset_world``): with
``np.random.Generator(PCG64(SeedSequence(seed)))`` -- ``gymnasium.utils.seeding``
-- a goal landmark ``choice`` followed by three agent positions and two
landmark positions ``uniform(-1, 1, 2)``.
"""
from __future__ import annotations

import numpy as np

AGENTS = ("adversary_0", "agent_0", "agent_1")
OBS_DIM = {"adversary_0": 8, "agent_0": 10, "agent_1": 10}
N_ACTIONS = 5
MAX_CYCLES = 25
ENV_SEED = 1870300          # utils/game_logic_functions.py:54
INIT_STATE_DIM = 11


class _Space:
    def __init__(self, shape=None, n=None):
        self.shape = shape
        self.n = n


class InitStateStream:
    """The PCG64 stream ``env.reset()`` consumes, one flat record per reset:
    [goal_idx, adversary.xy, agent_0.xy, agent_1.xy, landmark0.xy, landmark1.xy]."""

    _LANDMARKS = (object(), object())

    def __init__(self, seed=ENV_SEED):
        self.seed(seed)

    def seed(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def draw(self, n=1):
        out = np.empty((n, INIT_STATE_DIM), dtype=np.float64)
        rng = self.rng
        for i in range(n):
            # same call sequence as upstream reset_world
            goal = rng.choice(2)
            out[i, 0] = goal
            for a in range(3):
                out[i, 1 + 2 * a:3 + 2 * a] = rng.uniform(-1, +1, 2)
            for l in range(2):
                out[i, 7 + 2 * l:9 + 2 * l] = rng.uniform(-1, +1, 2)
        return out


class DeviceMPEEnv:
    """What ``initialize_env`` returns: the env *handle* of the device rollout.

    It exposes the attributes the reference reads from a PettingZoo env
    (``agents``, ``observation_space(a).shape``, ``action_space(a).n``,
    ``reset``, ``close``) and owns the initial-state stream.  Episodes are not
    stepped on the host: ``play_game`` and the training loops hand the pending
    initial states to the K1 kernels."""

    metadata = {"name": "simple_adversary_v3"}

    def __init__(self, render_mode=None, seed=ENV_SEED):
        if render_mode not in (None,):
            raise NotImplementedError("rendering is not part of the B200 hot path")
        self.possible_agents = list(AGENTS)
        self.agents = list(AGENTS)
        self.stream = InitStateStream(seed)
        self.pending = None

    def observation_space(self, agent):
        return _Space(shape=(OBS_DIM[agent],))

    def action_space(self, agent):
        return _Space(n=N_ACTIONS)

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.stream.seed(seed)
        self.pending = self.stream.draw(1)[0]

    def take_pending(self):
        """The initial state of the episode started by the last ``reset()``."""
        if self.pending is None:
            self.reset()
        rec, self.pending = self.pending, None
        return rec

    def draw_initial_states(self, n):
        """``n`` further resets' worth of initial states (batched evaluation)."""
        self.pending = None
        return self.stream.draw(n)

    def state_dict(self):
        """The PCG64 bit-generator state (checkpoint / resume keeps the reset stream in step)."""
        return {"bit_generator": self.stream.rng.bit_generator.state}

    def load_state_dict(self, sd):
        self.stream.rng.bit_generator.state = sd["bit_generator"]
        self.pending = None

    def close(self):
        pass
