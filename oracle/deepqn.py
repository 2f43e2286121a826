"""NumPy restatement of ``DeepQN.forward`` (K2's checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
/root/reference/Atari/deepqn.py:39-48 at the batch size the reference uses
(1): the "virtual batch norm" layers are plain ``BatchNorm2d`` in TRAIN mode
(SURVEY.md Appendix C #12), so statistics are per frame over H x W, biased
variance, eps 1e-5, affine gamma/beta.  Pinned against the reference module by
``oracle/make_golden.py``.
"""
from __future__ import annotations

import numpy as np

from . import layout

BN_EPS = np.float32(1e-5)


def _conv2d(x, w, b, stride):
    """x fp32[C,H,W], w fp32[O,C,kh,kw] -> fp32[O,Ho,Wo] (valid padding)."""
    C, H, W = x.shape
    O, _, kh, kw = w.shape
    Ho = (H - kh) // stride + 1
    Wo = (W - kw) // stride + 1
    s0, s1, s2 = x.strides
    cols = np.lib.stride_tricks.as_strided(
        x, shape=(Ho, Wo, C, kh, kw),
        strides=(s1 * stride, s2 * stride, s0, s1, s2), writeable=False)
    cols = cols.reshape(Ho * Wo, C * kh * kw).astype(np.float32)
    out = cols @ w.reshape(O, -1).T.astype(np.float32) + b.astype(np.float32)
    return out.T.reshape(O, Ho, Wo).astype(np.float32)


def _bn_train(x, g, b):
    """Per-frame train-mode BatchNorm2d: stats over H x W per channel."""
    mean = x.mean(axis=(1, 2), keepdims=True, dtype=np.float32)
    xc = x - mean
    var = (xc * xc).mean(axis=(1, 2), keepdims=True, dtype=np.float32)
    return xc / np.sqrt(var + BN_EPS) * g[:, None, None] + b[:, None, None]


def dqn_forward(row, frame, c_in, n_actions):
    """Logits fp32[A] for one frame (uint8 or float [C,84,84])."""
    p = layout.unpack_dqn(np.asarray(row, dtype=np.float32), c_in, n_actions)
    x = np.asarray(frame, dtype=np.float32) / np.float32(255)
    x = np.maximum(_bn_train(_conv2d(x, p["conv1.weight"], p["conv1.bias"], 4),
                             p["vbn1.weight"], p["vbn1.bias"]), 0)
    x = np.maximum(_bn_train(_conv2d(x, p["conv2.weight"], p["conv2.bias"], 2),
                             p["vbn2.weight"], p["vbn2.bias"]), 0)
    x = np.maximum(_bn_train(_conv2d(x, p["conv3.weight"], p["conv3.bias"], 1),
                             p["vbn3.weight"], p["vbn3.bias"]), 0)
    x = x.reshape(-1).astype(np.float32)
    x = np.maximum(p["fc1.weight"] @ x + p["fc1.bias"], 0)
    return (p["output.weight"] @ x + p["output.bias"]).astype(np.float32)


def dqn_forward_batch(rows, frames, c_in, n_actions):
    """rows fp32[P,D], frames [P,B,C,84,84] -> (logits fp32[P,B,A], actions int32[P,B])
    with the reference's first-maximum argmax (Atari/deepqn.py:55-60)."""
    P, B = frames.shape[:2]
    logits = np.zeros((P, B, n_actions), dtype=np.float32)
    for m in range(P):
        for f in range(B):
            logits[m, f] = dqn_forward(rows[m], frames[m, f], c_in, n_actions)
    return logits, np.argmax(logits, axis=-1).astype(np.int32)
