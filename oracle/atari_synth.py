"""CPU restatement of the synthetic Atari-like episode loop (N4's checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference's Atari rollout is dead code (``play_atari`` is called with 5
arguments but takes 4, ``DeepQN.forward(inputs, args)`` vs ``forward(self, x)``, ``len(actions)`` on a
``[1, A]`` tensor) and needs ROMs; there is nothing to run or to take golden vectors from.  This file
restates what that code evidently intends, with the three defects repaired and the emulator replaced by the
deterministic synthetic one of ``csrc/atari_synth.cu``:

* wrapper chain of ``initialize_env`` (/root/reference/utils/game_logic_functions.py:48-53):
  ``frame_stack_v1(env, 4)`` (oldest frame first, zeros before the first frames) and
  ``agent_indicator_v0`` (two one-hot planes; 255 on the acting agent's plane -- the upstream SuperSuit
  value is not verifiable here, stated as an assumption) => observations ``uint8[84, 84, 6]``;
* ``preprocess_observation`` (``:67-80``): HWC -> 1CHW float32;
* ``DeepQN.determine_action`` (/root/reference/Atari/deepqn.py:50-62): strict ``>`` scan = first maximum
  of the logits of ``forward`` (``:39-48``, oracle/deepqn.py);
* ``play_atari`` (``:84-119``): AEC order first_0, second_0; ``env.step(action)`` then ``env.last()``,
  whose reward is the NEXT agent's cumulative reward (PettingZoo's AEC bookkeeping, the same rotation as
  SURVEY.md Appendix B), credited to the agent that just acted; the loop stops at the agent-step limit.
"""
from __future__ import annotations

import numpy as np

from . import deepqn as odqn
from . import philox

FRAME = 84 * 84
BLOCKS = FRAME // 16


def _tag():
    return np.uint32(1 | (philox.KIND_FRAMES << 8))


def frame(seed, ep, t, j):
    """uint8[84, 84] frame of episode ``ep`` after ``t`` emulator steps, joint action ``j`` of step t."""
    k0, k1 = philox.split_seed(seed)
    b = np.arange(BLOCKS, dtype=np.uint32) | np.uint32(int(j) << 16)
    w = philox.philox4x32_10(b, np.uint32(ep), np.uint32(t), _tag(), k0, k1)
    words = np.stack(w, axis=-1).astype("<u4")                    # [441, 4] little-endian words
    return words.view(np.uint8).reshape(84, 84)


def reward_first(seed, ep, t, j):
    k0, k1 = philox.split_seed(seed)
    w = philox.philox4x32_10(np.uint32(0xFFFF | (int(j) << 16)), np.uint32(ep), np.uint32(t), _tag(), k0, k1)
    return np.float32((int(w[0]) & 0xFF) - 128) * np.float32(1.0 / 128.0)


def observe(frames, seat):
    """``frames``: list of the episode's frames so far (index = emulator step).  uint8[6, 84, 84]."""
    t = len(frames) - 1
    obs = np.zeros((6, 84, 84), dtype=np.uint8)
    for pl in range(4):
        ft = t - 3 + pl
        if ft >= 0:
            obs[pl] = frames[ft]
    obs[4 + seat] = 255
    return obs


def play_atari(row_first, row_second, n_actions, seed, ep, agent_step_limit, reference_compat=True,
               return_trace=False):
    """One episode; returns (rewards['first_0'], rewards['second_0']) like the repaired reference loop."""
    frames = [frame(seed, ep, 0, 0)]
    cum = {0: 0.0, 1: 0.0}                 # PettingZoo _cumulative_rewards
    got = {0: 0.0, 1: 0.0}                 # play_atari's rewards dict
    own = {0: 0.0, 1: 0.0}                 # each agent's own reward sum (reference_compat = False)
    steps, t, pending, trace = 0, 0, None, []
    while agent_step_limit is None or steps < agent_step_limit:
        seat = steps % 2
        obs = observe(frames, seat)
        logits = odqn.dqn_forward(row_first if seat == 0 else row_second, obs, 6, n_actions)
        action = int(np.argmax(logits))                        # first maximum
        trace.append((seat, action, logits))
        if seat == 0:
            pending = action
            rewards = (0.0, 0.0)                               # no emulator step yet: rewards cleared
        else:
            t += 1
            j = pending + 32 * action
            frames.append(frame(seed, ep, t, j))
            r = float(reward_first(seed, ep, t, j))
            rewards = (r, -r)
            own[0] += r
            own[1] += -r
        cum[seat] = 0.0
        cum[0] += rewards[0]
        cum[1] += rewards[1]
        got[seat] += cum[1 - seat]                             # env.last() reports the NEXT agent
        steps += 1
        if agent_step_limit is None and t >= 10000:
            break
    res = (got[0], got[1]) if reference_compat else (own[0], own[1])
    return (res, trace) if return_trace else res
