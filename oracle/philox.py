"""NumPy restatement of the counter-based noise the CUDA kernels regenerate.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference draws mutation noise from unseeded global generators
(``torch.normal`` in agent.py:25-29, ``np.random.normal`` in agent.py:51-53)
and is therefore irreproducible.  The B200 build
replaces both with Philox4x32-10 (Salmon et al., "Parallel random numbers: as
easy as 1, 2, 3", SC'11; Random123 known-answer vectors are checked in
``tests/test_philox.py``) keyed by the run seed, with the counter

    ctr = (param_index // 4, member, generation, role | kind << 8)

so any (member, parameter) normal can be regenerated anywhere without storing
noise.  Words -> normals by Box-Muller on word pairs (0,1) and (2,3).
Integer words are bit-exact between this file and the device; normals agree to
a few ulp (libm vs CUDA ``logf``/``sincospif``).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

KIND_ES = 0       # ES perturbation noise (agent.py:31-70 replacement)
KIND_GA = 1       # GA mutation noise     (agent.py:25-29 replacement)
KIND_ENV = 2      # device-side initial env states (Appendix A.3 replacement)
KIND_FRAMES = 3   # synthetic Atari frames
KIND_INIT = 4     # device-side founder initialisation
KIND_XOVER = 5    # GA crossover decisions / masks (extension; the reference has no crossover)

ROLE_ID = {"agent_0": 0, "agent_1": 1, "adversary_0": 2}


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u01(x):
    """uint32 -> float32 in (0, 1]: x * 2^-32 + 2^-33 (exact scaling)."""
    return (x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)


def box_muller(xa, xb):
    """Two uint32 words -> two float32 standard normals."""
    u1 = u01(xa)
    u2 = u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2.0) * u2).astype(np.float32)          # in (0, 2], units of pi
    a64 = ang.astype(np.float64) * np.pi
    return (r * np.cos(a64).astype(np.float32)).astype(np.float32), \
           (r * np.sin(a64).astype(np.float32)).astype(np.float32)


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)


def normals(seed, kind, role, gen, members, n_params):
    """float32[len(members), n_params] standard normals for flat parameter
    indices 0..n_params-1 of the given members (global member ids)."""
    k0, k1 = split_seed(seed)
    members = np.asarray(members, dtype=np.uint32).reshape(-1, 1)
    n4 = (n_params + 3) // 4
    j4 = np.arange(n4, dtype=np.uint32).reshape(1, -1)
    tag = np.uint32((int(role) & 0xFF) | (int(kind) << 8))
    x0, x1, x2, x3 = philox4x32_10(j4, members, np.uint32(gen), tag, k0, k1)
    z0, z1 = box_muller(x0, x1)
    z2, z3 = box_muller(x2, x3)
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(members.shape[0], n4 * 4)
    return z[:, :n_params]


def words(seed, kind, role, gen, members, n4):
    """Raw uint32[len(members), n4, 4] Philox words (bit-exact device check)."""
    k0, k1 = split_seed(seed)
    members = np.asarray(members, dtype=np.uint32).reshape(-1, 1)
    j4 = np.arange(n4, dtype=np.uint32).reshape(1, -1)
    tag = np.uint32((int(role) & 0xFF) | (int(kind) << 8))
    return np.stack(philox4x32_10(j4, members, np.uint32(gen), tag, k0, k1), axis=-1)


def init_states(seed, stream_id, n, rec0=0):
    """fp64[n,11] synthetic initial env states, bit-exact restatement of the
    device generator (``cev_init_states_f64``): goal = word0 & 1, the ten
    coordinates U(-1,1) from words 1..10 of three Philox blocks per record."""
    k0, k1 = split_seed(seed)
    rec = np.arange(rec0, rec0 + n, dtype=np.uint32)
    tag = np.uint32(KIND_ENV << 8)
    w = []
    for b in range(3):
        w.extend(philox4x32_10(np.uint32(b), rec, np.uint32(stream_id), tag, k0, k1))
    out = np.zeros((n, 11))
    out[:, 0] = (w[0] & np.uint32(1)).astype(np.float64)
    for i in range(10):
        out[:, 1 + i] = (w[1 + i].astype(np.float64) + 0.5) * (2.0 / 4294967296.0) - 1.0
    return out


def fc_init_rows(seed, role, members, in_dim):
    """fp32[len(members), D] founders, bit-exact restatement of ``cev_fc_init_f32``:
    Linear ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from one Philox word per parameter,
    LayerNorm gamma = 1, beta = 0 (PyTorch default init of MPE/fcnetwork.py:11-22)."""
    from . import layout
    segs, total = layout.fc_segments(in_dim)
    w = words(seed, KIND_INIT, role, 0, members, (total + 3) // 4).reshape(len(members), -1)[:, :total]
    u = (w.astype(np.float64) + 0.5) * (2.0 / 4294967296.0) - 1.0
    out = np.zeros((len(members), total), dtype=np.float32)
    fan = {"fc1": in_dim, "fc2": layout.H1, "output": layout.H2}
    for name, off, shape, _ in segs:
        n = int(np.prod(shape))
        mod, kind = name.split(".")
        if mod.startswith("ln"):
            out[:, off:off + n] = 1.0 if kind == "weight" else 0.0
        else:
            out[:, off:off + n] = (u[:, off:off + n] * (1.0 / np.sqrt(float(fan[mod])))).astype(np.float32)
    return out
