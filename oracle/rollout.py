"""Vectorised CPU restatement of the reference MPE rollout (K1's checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows, for a batch of independent episodes:
* ``FCNetwork.forward``            /root/reference/MPE/fcnetwork.py:37-70
* ``FCNetwork.determine_action``   /root/reference/MPE/fcnetwork.py:73-90
  (strict ``>`` scan from index 0 => lowest index wins ties)
* ``play_MPE``                     /root/reference/utils/game_logic_functions.py:123-212
  (reward attribution rotated by one seat)
* the environment of ``oracle/mpe_env.py`` (fp64 physics, fp32 observations)

The network runs in fp32 (NumPy), the environment in fp64, like the reference.
Pinned against the reference's own ``play_game`` by ``oracle/make_golden.py``
/ ``tests/test_oracle_golden.py`` (the reference-generated fixtures under ``tests/golden/``).
"""
from __future__ import annotations

import numpy as np

from . import layout, mpe_env

LN_EPS = np.float32(1e-5)
SEATS = ("adversary_0", "agent_0", "agent_1")   # world / AEC order


def layer_norm(x, g, b):
    """Biased-variance LayerNorm over the last axis in fp32
    (``nn.LayerNorm`` defaults, MPE/fcnetwork.py:16,19)."""
    x = x.astype(np.float32, copy=False)
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    xc = x - mean
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=np.float32)
    rstd = np.float32(1.0) / np.sqrt(var + LN_EPS)
    return xc * rstd * g + b


def fc_forward(row, obs, in_dim):
    """Logits fp32[B,5] of one flat network ``row`` on ``obs`` fp32[B,in]."""
    p = layout.unpack_fc(np.asarray(row, dtype=np.float32), in_dim)
    x = np.asarray(obs, dtype=np.float32)
    h = x @ p["fc1.weight"].T + p["fc1.bias"]
    h = np.maximum(layer_norm(h, p["ln1.weight"], p["ln1.bias"]), np.float32(0))
    h = h @ p["fc2.weight"].T + p["fc2.bias"]
    h = np.maximum(layer_norm(h, p["ln2.weight"], p["ln2.bias"]), np.float32(0))
    return (h @ p["output.weight"].T + p["output.bias"]).astype(np.float32)


def argmax_first(logits):
    """Strict-``>`` scan from index 0 (lowest index on ties); also returns the
    top-2 gap used as the decision margin in parity tests."""
    logits = np.asarray(logits)
    act = np.argmax(logits, axis=-1)          # numpy argmax returns first max
    srt = np.sort(logits, axis=-1)
    gap = srt[..., -1] - srt[..., -2]
    return act.astype(np.int32), gap.astype(np.float32)


_ACT_U = np.array([[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1]], dtype=np.float64)


def observations(pos, lm, goal):
    """fp32 observations for the three seats from fp64 state.
    pos[N,3,2], lm[N,2,2], goal[N] -> (adv[N,8], a0[N,10], a1[N,10])."""
    n = pos.shape[0]
    gpos = lm[np.arange(n), goal]                      # [N,2]
    out = []
    for i in range(3):
        me = pos[:, i]
        ent = [lm[:, 0] - me, lm[:, 1] - me]
        oth = [pos[:, j] - me for j in range(3) if j != i]
        parts = ent + oth if i == 0 else [gpos - me] + ent + oth
        out.append(np.concatenate(parts, axis=1).astype(np.float32))
    return out


def world_step(pos, vel, actions, pos_first=True):
    """One physics step in fp64 (Appendix A.2/A.4); in place."""
    u = _ACT_U[actions] * mpe_env.SENSITIVITY          # [N,3,2]
    if pos_first:
        pos += vel * mpe_env.DT
    vel *= (1 - mpe_env.DAMPING)
    vel += (u / mpe_env.MASS) * mpe_env.DT
    if not pos_first:
        pos += vel * mpe_env.DT


def step_rewards(pos, lm, goal):
    """(r_good[N], r_adv[N]) after a world step (Appendix A.5)."""
    n = pos.shape[0]
    g = lm[np.arange(n), goal]
    d = np.sqrt(np.sum(np.square(pos - g[:, None, :]), axis=2))   # [N,3]
    r_adv = -d[:, 0]
    r_good = -np.minimum(d[:, 1], d[:, 2]) + d[:, 0]
    return r_good, r_adv


def rollout(nets, idx, init, n_cycles=25, pos_first=True, forced_actions=None,
            return_traces=False):
    """Play ``N`` independent episodes.

    nets : dict seat -> fp32[n_seat, >=D_seat] flat rows (``parameters()`` order)
    idx  : int[N,3]  row index per seat in SEATS order (adversary_0, agent_0, agent_1)
    init : fp64[N,11] flat initial-state records (``mpe_env.flat_state``)
    forced_actions : optional int[N,n_cycles,3]; replay these actions instead
        of the argmax (teacher forcing) while still reporting logits

    Returns dict with fp64 ``sum_good`` (sum_c r_good), ``last_good`` (r_good of
    the final cycle), ``sum_adv``, fp32 ``min_gap`` (smallest top-2 logit gap of
    any decision in the episode), int32 ``actions[N,n_cycles,3]`` and, if
    requested, ``logits[N,n_cycles,3,5]``.
    """
    idx = np.asarray(idx, dtype=np.int64)
    init = np.asarray(init, dtype=np.float64)
    n = idx.shape[0]
    goal = init[:, 0].astype(np.int64)
    pos = init[:, 1:7].reshape(n, 3, 2).copy()
    vel = np.zeros_like(pos)
    lm = init[:, 7:11].reshape(n, 2, 2).copy()
    sum_good = np.zeros(n)
    last_good = np.zeros(n)
    sum_adv = np.zeros(n)
    min_gap = np.full(n, np.inf, dtype=np.float32)
    actions = np.zeros((n, n_cycles, 3), dtype=np.int32)
    logits_tr = np.zeros((n, n_cycles, 3, 5), dtype=np.float32) if return_traces else None
    groups = []
    for s, seat in enumerate(SEATS):
        g = {}
        for e in range(n):
            g.setdefault(int(idx[e, s]), []).append(e)
        groups.append({k: np.asarray(v) for k, v in g.items()})
    for c in range(n_cycles):
        obs = observations(pos, lm, goal)
        for s, seat in enumerate(SEATS):
            in_dim = layout.OBS_DIM[seat]
            for row, eps in groups[s].items():
                lg = fc_forward(nets[seat][row], obs[s][eps], in_dim)
                if not np.all(np.isfinite(lg)):
                    raise ValueError("\n\t Warning: output contains inf or NaN")
                a, gap = argmax_first(lg)
                actions[eps, c, s] = a
                min_gap[eps] = np.minimum(min_gap[eps], gap)
                if return_traces:
                    logits_tr[eps, c, s] = lg
        act = actions[:, c] if forced_actions is None else np.asarray(forced_actions)[:, c]
        if forced_actions is not None:
            actions[:, c] = act
        world_step(pos, vel, act, pos_first)
        rg, ra = step_rewards(pos, lm, goal)
        sum_good += rg
        sum_adv += ra
        last_good = rg
    out = dict(sum_good=sum_good, last_good=np.asarray(last_good, dtype=np.float64).copy(),
               sum_adv=sum_adv, min_gap=min_gap, actions=actions)
    if return_traces:
        out["logits"] = logits_tr
    return out


def cycles_for_limit(agent_step_limit):
    """Number of world steps the reference executes under an agent-step limit
    (``play_MPE`` counts AEC agent-steps, game_logic_functions.py:127,195)."""
    if agent_step_limit is None:
        return mpe_env.MAX_CYCLES
    return int(min(mpe_env.MAX_CYCLES, max(0, agent_step_limit) // 3))


def compat_slots(res, agent_step_limit=None):
    """Map physical reward sums to what ``play_MPE`` returns
    ``(rewards['agent_0'], rewards['agent_1'], rewards['adversary_0'])`` under
    the rotated attribution (Appendix B).  ``res`` must come from a rollout of
    ``cycles_for_limit(limit)`` cycles."""
    L = 75 if agent_step_limit is None else int(min(75, max(0, agent_step_limit)))
    nc = L // 3
    n_adv = -(-L // 3)              # adversary turns taken
    n_a0 = -(-(L - 1) // 3) if L >= 1 else 0
    sg, lg, sa = res["sum_good"], res["last_good"], res["sum_adv"]

    def good_prefix(k):            # sum_{c<=k} r_good(c) for k in {nc, nc-1}
        if k >= nc:
            return sg
        return sg - lg
    slot_adv = good_prefix(max(n_adv - 1, 0)) if nc > 0 else np.zeros_like(sg)
    slot_a0 = good_prefix(max(n_a0 - 1, 0)) if nc > 0 else np.zeros_like(sg)
    slot_a1 = sa
    return slot_a0, slot_a1, slot_adv


def true_slots(res):
    """Each role's own reward sum (the ``reference_compat=False`` mode)."""
    return res["sum_good"], res["sum_good"], res["sum_adv"]
