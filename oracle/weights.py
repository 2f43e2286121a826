"""Deterministic synthetic weights for tests and benches (NumPy PCG64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Draws from the same
distribution as the reference's default PyTorch initialisation
(``create_agent`` -> ``MPEAgent`` -> ``FCNetwork``, MPE/fcnetwork.py:11-22:
``nn.Linear`` kaiming-uniform(a=sqrt(5)) => U(-1/sqrt(fan_in), 1/sqrt(fan_in))
for weight and bias; LayerNorm gamma=1, beta=0) but from NumPy's PCG64 so the
vectors are reproducible on any box without shipping megabytes of fixtures.
"""
from __future__ import annotations

import numpy as np

from . import layout


def make_fc_rows(n, in_dim, seed, ln_jitter=0.0):
    """fp32[n, D(in_dim)] flat rows in ``parameters()`` order.

    ``ln_jitter`` > 0 perturbs LayerNorm gamma/beta (GA mutates them,
    agent.py:25-29) so tests exercise the affine part."""
    rng = np.random.Generator(np.random.PCG64(seed))
    segs, total = layout.fc_segments(in_dim)
    rows = np.zeros((n, total), dtype=np.float32)
    fan_in = {"fc1": in_dim, "fc2": layout.H1, "output": layout.H2}
    for name, off, shape, _ in segs:
        size = int(np.prod(shape))
        mod, kind = name.split(".")
        if mod.startswith("ln"):
            base = 1.0 if kind == "weight" else 0.0
            v = base + ln_jitter * rng.standard_normal((n, size))
        else:
            bound = 1.0 / np.sqrt(fan_in[mod])
            v = rng.uniform(-bound, bound, (n, size))
        rows[:, off:off + size] = v.astype(np.float32)
    return rows


def make_dqn_rows(n, c_in, n_actions, seed, bn_jitter=0.0):
    """fp32[n, D] DeepQN rows (Atari/deepqn.py:7-36 default init)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    segs, total = layout.dqn_segments(c_in, n_actions)
    rows = np.zeros((n, total), dtype=np.float32)
    fan_in = {"conv1": c_in * 64, "conv2": 32 * 16, "conv3": 64 * 9,
              "fc1": 3136, "output": 512}
    for name, off, shape, _ in segs:
        size = int(np.prod(shape))
        mod, kind = name.split(".")
        if mod.startswith("vbn"):
            base = 1.0 if kind == "weight" else 0.0
            v = base + bn_jitter * rng.standard_normal((n, size))
        else:
            bound = 1.0 / np.sqrt(fan_in[mod])
            v = rng.uniform(-bound, bound, (n, size))
        rows[:, off:off + size] = v.astype(np.float32)
    return rows
