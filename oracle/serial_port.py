"""Episode-serial CPU port that executes the way the reference does.

TEST INFRASTRUCTURE / CPU BASELINE (see ``oracle/__init__.py``).  The GPU box
has no ``/root/reference``, so ``bench.py``'s CPU-baseline legs time this port:
one episode at a time, one agent-step at a time, batch-1 ``torch`` forward with
the reference's per-call NaN/Inf scans and Python arg-max loop
(MPE/fcnetwork.py:37-90), driven by the AEC loop of
utils/game_logic_functions.py:123-212 on ``oracle.mpe_env``.  It is checked
against the vectorised oracle in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layout, mpe_env


class PortFCNetwork(nn.Module):
    """Restatement of the reference policy network (MPE/fcnetwork.py:9-70)."""

    def __init__(self, in_dim):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, layout.H1)
        self.ln1 = nn.LayerNorm(layout.H1)
        self.fc2 = nn.Linear(layout.H1, layout.H2)
        self.ln2 = nn.LayerNorm(layout.H2)
        self.output = nn.Linear(layout.H2, layout.NACT)

    @staticmethod
    def _guard(x, where):
        if torch.isinf(x).any() or torch.isnan(x).any():
            raise ValueError(f"\n\t Warning: output contains inf or NaN {where}")

    def forward(self, x):
        self._guard(x, "(input)")
        x = F.relu(self.ln1(self.fc1(x)))
        self._guard(x, "after fc1")
        x = F.relu(self.ln2(self.fc2(x)))
        self._guard(x, "after fc2")
        x = self.output(x)
        self._guard(x, "")
        return x

    def determine_action(self, obs):
        logits = self.forward(obs)
        best, pos = -float("inf"), -1
        for i in range(len(logits)):
            if logits[i] > best:
                pos, best = i, logits[i]
        return pos

    @classmethod
    def from_row(cls, row, in_dim):
        net = cls(in_dim)
        p = layout.unpack_fc(np.asarray(row, dtype=np.float32), in_dim)
        net.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in p.items()})
        return net


def play_episode(env, nets, agent_step_limit=None):
    """One episode on ``env`` (already reset).  ``nets``: dict seat name ->
    PortFCNetwork.  Returns the reference's (agent_0, agent_1, adversary_0) triple
    with its rotated attribution (acting agent is credited ``env.last()``)."""
    rewards = {"agent_0": 0.0, "agent_1": 0.0, "adversary_0": 0.0}
    steps = 0
    with torch.no_grad():
        for agent in env.agent_iter():
            obs = torch.from_numpy(env.observe(agent)).to(torch.float32)
            action = nets[agent].determine_action(obs)
            env.step(action)
            _, reward, term, trunc, _ = env.last()
            rewards[agent] += reward
            steps += 1
            if agent_step_limit is not None and steps >= agent_step_limit:
                break
            if term or trunc:
                break
    return rewards["agent_0"], rewards["agent_1"], rewards["adversary_0"]


def play_rows(rows, init_states, agent_step_limit=None):
    """Episodes for explicit initial states; ``rows``: dict seat -> flat fp32 row."""
    torch.set_num_threads(1)
    nets = {s: PortFCNetwork.from_row(rows[s], layout.OBS_DIM[s]) for s in rows}
    env = mpe_env.SimpleAdversaryEnv()
    out = np.zeros((len(init_states), 3))
    for i, rec in enumerate(init_states):
        env.load_flat_state(rec)
        out[i] = play_episode(env, nets, agent_step_limit)
    return out


def _worker(job):
    rows, init_states = job
    import time
    t0 = time.perf_counter()
    out = play_rows(rows, init_states)
    return out, time.perf_counter() - t0


def timed_sample(rows, init_states, n_procs=1):
    """Run the sample over ``n_procs`` independent processes (the reference itself
    is serial; >1 is the embarrassingly-parallel upper bound).  Returns
    (rewards[N,3], wall seconds)."""
    import time
    if n_procs <= 1:
        t0 = time.perf_counter()
        out = play_rows(rows, init_states)
        return out, time.perf_counter() - t0
    import multiprocessing as mp
    chunks = [c for c in np.array_split(np.arange(len(init_states)), n_procs) if len(c)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_worker, [(rows, init_states[c]) for c in chunks])
    wall = time.perf_counter() - t0
    return np.concatenate([r[0] for r in res]), wall
