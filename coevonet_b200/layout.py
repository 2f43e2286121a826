"""Flat parameter-row layouts shared by the Python boundary and the kernels.

Device populations are ``float32[P, pitch]`` tensors whose rows hold a network
in the reference's ``parameters()`` order (``MPE/fcnetwork.py:11-22``,
SURVEY.md Appendix D) with the row pitch padded to a multiple of 32 floats so
every row and every sub-tensor used by ``cp.async``/128-bit loads is 16-byte
aligned.  The perturbable (ES) view skips the LayerNorm segments
(``MPE/fcnetwork.py:185-199``).
"""
from __future__ import annotations

import numpy as np
import torch

H1, H2, NACT = 512, 256, 5
ROLES = ("agent_0", "agent_1", "adversary_0")          # reference return order
SEATS = ("adversary_0", "agent_0", "agent_1")          # world / AEC order
SEAT_OF = {name: i for i, name in enumerate(SEATS)}
OBS_DIM = {"adversary_0": 8, "agent_0": 10, "agent_1": 10}
ROLE_ID = {"agent_0": 0, "agent_1": 1, "adversary_0": 2}   # Philox role tag


def _round_up(x, m):
    return (x + m - 1) // m * m


def fc_segments(in_dim):
    """[(state_dict name, offset, shape, perturbable)] in ``parameters()`` order."""
    spec = (("fc1.weight", (H1, in_dim), True), ("fc1.bias", (H1,), True),
            ("ln1.weight", (H1,), False), ("ln1.bias", (H1,), False),
            ("fc2.weight", (H2, H1), True), ("fc2.bias", (H2,), True),
            ("ln2.weight", (H2,), False), ("ln2.bias", (H2,), False),
            ("output.weight", (NACT, H2), True), ("output.bias", (NACT,), True))
    segs, off = [], 0
    for name, shape, pert in spec:
        segs.append((name, off, shape, pert))
        off += int(np.prod(shape))
    return segs, off


def fc_dim(in_dim):
    return fc_segments(in_dim)[1]


def fc_pitch(in_dim):
    return _round_up(fc_dim(in_dim), 32)


def fc_perturbable_index(in_dim):
    """int64 indices of the Linear parameters inside a flat row, in the order of
    ``get_perturbable_weights`` (``MPE/fcnetwork.py:155-199``)."""
    segs, _ = fc_segments(in_dim)
    return np.concatenate([np.arange(off, off + int(np.prod(shape)), dtype=np.int64)
                           for _, off, shape, pert in segs if pert])


def pack_state_dict(sd, in_dim, out=None):
    """state_dict -> float32[pitch] CPU tensor (padding zero)."""
    segs, total = fc_segments(in_dim)
    row = torch.zeros(fc_pitch(in_dim), dtype=torch.float32) if out is None else out
    for name, off, shape, _ in segs:
        row[off:off + int(np.prod(shape))] = sd[name].detach().to(torch.float32).reshape(-1).cpu()
    return row


def unpack_to_state_dict(row, in_dim):
    """flat row (any device) -> dict of CPU float32 tensors keyed like state_dict."""
    segs, _ = fc_segments(in_dim)
    row = row.detach().to("cpu", torch.float32)
    return {name: row[off:off + int(np.prod(shape))].reshape(shape).clone()
            for name, off, shape, _ in segs}


def pack_models(models, in_dim, device=None):
    """list of FCNetwork-like modules -> float32[len, pitch] tensor."""
    rows = torch.zeros((len(models), fc_pitch(in_dim)), dtype=torch.float32)
    for i, m in enumerate(models):
        pack_state_dict(m.state_dict(), in_dim, out=rows[i])
    return rows.to(device) if device is not None else rows


# --- DeepQN (Atari/deepqn.py:7-36) -----------------------------------------
def dqn_segments(c_in, n_actions):
    spec = (("conv1.weight", (32, c_in, 8, 8), True), ("conv1.bias", (32,), True),
            ("conv2.weight", (64, 32, 4, 4), True), ("conv2.bias", (64,), True),
            ("conv3.weight", (64, 64, 3, 3), True), ("conv3.bias", (64,), True),
            ("fc1.weight", (512, 3136), True), ("fc1.bias", (512,), True),
            ("output.weight", (n_actions, 512), True), ("output.bias", (n_actions,), True),
            ("vbn1.weight", (32,), False), ("vbn1.bias", (32,), False),
            ("vbn2.weight", (64,), False), ("vbn2.bias", (64,), False),
            ("vbn3.weight", (64,), False), ("vbn3.bias", (64,), False))
    segs, off = [], 0
    for name, shape, pert in spec:
        segs.append((name, off, shape, pert))
        off += int(np.prod(shape))
    return segs, off


def dqn_dim(c_in, n_actions):
    return dqn_segments(c_in, n_actions)[1]


def dqn_pitch(c_in, n_actions):
    return _round_up(dqn_dim(c_in, n_actions), 32)


def pack_dqn_state_dict(sd, c_in, n_actions):
    segs, _ = dqn_segments(c_in, n_actions)
    row = torch.zeros(dqn_pitch(c_in, n_actions), dtype=torch.float32)
    for name, off, shape, _ in segs:
        row[off:off + int(np.prod(shape))] = sd[name].detach().to(torch.float32).reshape(-1).cpu()
    return row
