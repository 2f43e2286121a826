"""Flat parameter layouts of the reference networks (oracle-side restatement).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Independent restatement of
SURVEY.md Appendix D; ``tests/test_layout.py`` checks it against both the
reference modules and the product's ``coevonet_b200.layout``.

``parameters()`` order of ``FCNetwork`` (``MPE/fcnetwork.py:11-22``):
fc1.W[512,in] fc1.b ln1.g ln1.b fc2.W[256,512] fc2.b ln2.g ln2.b out.W[5,256] out.b
"""
from __future__ import annotations

import numpy as np

H1, H2, NACT = 512, 256, 5
OBS_DIM = {"adversary_0": 8, "agent_0": 10, "agent_1": 10}
ROLES = ("agent_0", "agent_1", "adversary_0")


def fc_segments(in_dim):
    """[(name, offset, shape, perturbable)] in ``parameters()`` order."""
    segs = []
    off = 0
    for name, shape, pert in (
        ("fc1.weight", (H1, in_dim), True), ("fc1.bias", (H1,), True),
        ("ln1.weight", (H1,), False), ("ln1.bias", (H1,), False),
        ("fc2.weight", (H2, H1), True), ("fc2.bias", (H2,), True),
        ("ln2.weight", (H2,), False), ("ln2.bias", (H2,), False),
        ("output.weight", (NACT, H2), True), ("output.bias", (NACT,), True),
    ):
        segs.append((name, off, shape, pert))
        off += int(np.prod(shape))
    return segs, off


def fc_dim(in_dim):
    return fc_segments(in_dim)[1]


def fc_perturbable_index(in_dim):
    """int64 indices (into the full flat row) of the perturbable parameters,
    in ``get_perturbable_weights`` order (``MPE/fcnetwork.py:155-199``)."""
    segs, _ = fc_segments(in_dim)
    return np.concatenate([np.arange(off, off + int(np.prod(shape)))
                           for _, off, shape, pert in segs if pert])


def unpack_fc(row, in_dim):
    """flat fp32 row -> dict of arrays (views) keyed by state_dict names."""
    segs, total = fc_segments(in_dim)
    row = np.asarray(row)
    assert row.shape[-1] >= total
    return {name: row[off:off + int(np.prod(shape))].reshape(shape)
            for name, off, shape, _ in segs}


def pack_fc_state_dict(sd, in_dim):
    """torch state_dict (or dict of arrays) -> flat fp32 row."""
    segs, total = fc_segments(in_dim)
    out = np.empty(total, dtype=np.float32)
    for name, off, shape, _ in segs:
        v = sd[name]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        out[off:off + v.size] = v.reshape(-1)
    return out


# --- DeepQN (Atari/deepqn.py:7-36) ------------------------------------------
def dqn_segments(c_in, n_actions):
    segs = []
    off = 0
    for name, shape, pert in (
        ("conv1.weight", (32, c_in, 8, 8), True), ("conv1.bias", (32,), True),
        ("conv2.weight", (64, 32, 4, 4), True), ("conv2.bias", (64,), True),
        ("conv3.weight", (64, 64, 3, 3), True), ("conv3.bias", (64,), True),
        ("fc1.weight", (512, 3136), True), ("fc1.bias", (512,), True),
        ("output.weight", (n_actions, 512), True), ("output.bias", (n_actions,), True),
        ("vbn1.weight", (32,), False), ("vbn1.bias", (32,), False),
        ("vbn2.weight", (64,), False), ("vbn2.bias", (64,), False),
        ("vbn3.weight", (64,), False), ("vbn3.bias", (64,), False),
    ):
        segs.append((name, off, shape, pert))
        off += int(np.prod(shape))
    return segs, off


def unpack_dqn(row, c_in, n_actions):
    segs, total = dqn_segments(c_in, n_actions)
    row = np.asarray(row)
    return {name: row[off:off + int(np.prod(shape))].reshape(shape)
            for name, off, shape, _ in segs}


def pack_dqn_state_dict(sd, c_in, n_actions):
    segs, total = dqn_segments(c_in, n_actions)
    out = np.empty(total, dtype=np.float32)
    for name, off, shape, _ in segs:
        v = sd[name]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        out[off:off + v.size] = v.reshape(-1)
    return out
