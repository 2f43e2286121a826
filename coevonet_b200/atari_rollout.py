"""Batched synthetic Atari-like rollout around the K2 forward.

The reference's Atari episode loop (``play_atari``, ``utils/game_logic_functions.py:84-119``) is dead code
and needs ROMs; this is the device form of what it intends, with the emulator
replaced by the deterministic synthetic one of ``csrc/atari_synth.cu``: ``P`` population members in one
seat play ``E`` episodes each against one opponent row in the other seat.  Per AEC cycle: build both agents'
observations from the frame ring (``frame_stack_v1(4)`` + ``agent_indicator_v0``: C = 6), one grouped K2
forward for the members (their own weights, E frames each) and one for the opponent (one row, P*E frames),
one emulator step.  Rewards follow ``env.last()`` after ``env.step()`` (the next agent's cumulative reward,
credited to the agent that acted) under ``reference_compat``.
"""
from __future__ import annotations

import torch

from . import layout, ops


def cycles_for_limit(agent_step_limit):
    """(complete emulator steps, whether first_0 acts once more) under an agent-step limit."""
    return int(agent_step_limit) // 2, int(agent_step_limit) % 2 == 1


def atari_rollout(member_seat, members, opponent, n_actions, n_envs, agent_step_limit, seed, *, ep0=0,
                  reference_compat=True):
    """members fp32[P, pitch] (DeepQN rows, C = 6), opponent fp32[pitch].  Returns
    (reward_first, reward_second) fp64[P, E] as the repaired ``play_atari`` returns them."""
    if members.device.type != "cuda":
        raise RuntimeError("atari_rollout needs CUDA tensors (no CPU fallback)")
    dev = members.device
    P, E = members.shape[0], int(n_envs)
    N = P * E
    c_in = 6
    if members.stride(0) != layout.dqn_pitch(c_in, n_actions):
        raise ValueError("members must be DeepQN(6, n_actions) rows with the padded pitch")
    ring = torch.zeros((N, 4, 84 * 84), dtype=torch.uint8, device=dev)
    ops.atari_synth_step(seed, ep0, ring, 0)
    n_cycles, odd = cycles_for_limit(agent_step_limit)
    opp = opponent.reshape(1, -1).contiguous()
    r_hist = []

    def act(seat, t):
        obs = ops.atari_observe(ring, t, seat)
        if seat == member_seat:
            _, a = ops.deepqn_forward(members, obs.reshape(P, E, c_in, 84, 84), c_in, n_actions)
        else:
            _, a = ops.deepqn_forward(opp, obs.reshape(1, N, c_in, 84, 84), c_in, n_actions)
        return a.reshape(N).contiguous()

    for t in range(1, n_cycles + 1):
        a_first = act(0, t - 1)
        a_second = act(1, t - 1)
        r_hist.append(ops.atari_synth_step(seed, ep0, ring, t, a_first, a_second).to(torch.float64))
    if odd:
        act(0, n_cycles)                      # first_0's last turn: its action never reaches the emulator
    if not r_hist:
        z = torch.zeros((P, E), dtype=torch.float64, device=dev)
        return z, z.clone()
    r = torch.stack(r_hist, dim=0)            # [cycles, N] r_first per emulator step; r_second = -r_first
    if reference_compat:
        # first_0's turn in cycle c reads second_0's reward of cycle c-1; second_0's turn reads first_0's of cycle c
        n_first_turns = n_cycles + (1 if odd else 0)
        got_first = -(r[:max(n_first_turns - 1, 0)].sum(dim=0))
        got_second = r.sum(dim=0)
    else:
        got_first, got_second = r.sum(dim=0), -r.sum(dim=0)
    return got_first.reshape(P, E), got_second.reshape(P, E)
