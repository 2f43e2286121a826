"""The two arithmetic devices of the lockstep opponent kernel (coevonet_b200/csrc/rollout_lockstep.cu),
restated in NumPy: (1) LayerNorm-1 statistics of y = W1 x + b1 in closed form from the row covariance of
[W1 | b1]; (2) the 3xTF32 product (hi.hi + lo.hi + hi.lo with operands read truncated to TF32).  Both must
stay at fp32 rounding level, because the policy is an argmax and parity is asserted on decision margins of
1e-4.  CPU only."""
import numpy as np

from oracle import weights


def _tf32_trunc(x):
    """What the tensor core reads of an fp32 word: the low 13 mantissa bits dropped."""
    return (np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away, on the magnitude bits."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_closed_form_layernorm_statistics_match_two_pass():
    rng = np.random.default_rng(5)
    for in_dim, seed in ((10, 1), (8, 3)):
        row = weights.make_fc_rows(1, in_dim, seed)[0]
        w1 = row[:512 * in_dim].reshape(512, in_dim).astype(np.float64)
        b1 = row[512 * in_dim:512 * in_dim + 512].astype(np.float64)
        aug = np.concatenate([w1, b1[:, None]], axis=1)               # [512, in+1]
        wbar = aug.mean(axis=0)
        cov = (aug - wbar).T @ (aug - wbar) / 512.0                   # biased, like nn.LayerNorm
        for _ in range(200):
            x = rng.uniform(-2.5, 2.5, in_dim).astype(np.float32).astype(np.float64)
            z = np.concatenate([x, [1.0]])
            y = w1 @ x + b1
            assert abs(wbar @ z - y.mean()) <= 1e-12 * max(1.0, abs(y.mean()))
            assert abs(z @ cov @ z - y.var()) <= 1e-11 * y.var()
            # and at fp32 level against the reference's fp32 two-pass statistics
            y32 = (w1.astype(np.float32) @ x.astype(np.float32) + b1.astype(np.float32)).astype(np.float32)
            m32 = y32.mean(dtype=np.float32)
            v32 = np.mean((y32 - m32) ** 2, dtype=np.float32)
            assert abs(np.float32(wbar @ z) - m32) <= 4e-7 * max(1.0, abs(m32))
            assert abs(np.float32(z @ cov @ z) - v32) <= 2e-6 * v32


def test_three_tf32_products_reach_fp32_accuracy():
    rng = np.random.default_rng(7)
    K, rows, cols = 512, 64, 48
    h = np.maximum(rng.normal(0.3, 1.0, (rows, K)), 0).astype(np.float32)        # post-ReLU activations
    w = rng.uniform(-1 / np.sqrt(K), 1 / np.sqrt(K), (cols, K)).astype(np.float32)
    a_hi = _tf32_trunc(h)                       # producers: hi = h & 0xffffe000, lo = h - hi (exact)
    a_lo = (h - a_hi).astype(np.float32)
    assert np.array_equal(a_hi.astype(np.float64) + a_lo.astype(np.float64), h.astype(np.float64))
    b_hi = _tf32_rna(w)                         # ls_split_w2_kernel: cvt.rna hi, cvt.rna of the remainder
    b_lo = _tf32_rna((w - b_hi).astype(np.float32))
    rd = lambda t: _tf32_trunc(t).astype(np.float64)          # the operand as the tensor core reads it
    got = rd(a_lo) @ rd(b_hi).T + rd(a_hi) @ rd(b_lo).T + rd(a_hi) @ rd(b_hi).T
    exact = h.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(h).astype(np.float64) @ np.abs(w).astype(np.float64).T
    assert np.max(np.abs(got - exact) / scale) < 2.0 ** -20
    # one TF32 pass alone is three orders of magnitude worse: that is why the kernel issues three MMAs
    single = rd(h) @ rd(w).T
    assert np.max(np.abs(single - exact) / scale) > 2.0 ** -13


def _scale_exp(bound, target):
    """ls_scale_exp (rollout_lockstep.cu): 2^s * bound < 2^(target + 1), clamped to [-40, 40]."""
    if not (bound > 0) or not np.isfinite(bound):
        return 0
    return int(np.clip(target - int(np.floor(np.log2(np.float32(bound)))), -40, 40))


def _fp16_split(v):
    """v' -> (fp16(v'), fp16(v' - fp16(v'))), as the opponent kernel's producers and ls_split_w2_kernel do."""
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float32).astype(np.float16)
    return hi, lo


def test_scaled_two_term_fp16_products_reach_fp32_accuracy():
    """The opponent kernel since round 2: x1.w1 + x2.w1 + x1.w2 with power-of-two scaled FP16 hi / lo parts
    (kind::f16 MMAs, fp32 accumulation, the accumulator unscaled by an exact power of two).  Checked over weight
    and LayerNorm-parameter magnitudes far outside what training produces: the scales must keep every hi part
    finite and the product at fp32 rounding level."""
    rng = np.random.default_rng(11)
    K, rows, cols = 512, 64, 48
    for w_mag, gamma, beta in ((1 / np.sqrt(K), 1.0, 0.0), (300.0, 1.0, 0.0), (1e-6, 1.0, 0.0), (0.05, 80.0, 5.0),
                               (0.05, 1e-3, 0.0), (2e4, 40.0, 100.0)):
        # post-ReLU LayerNorm outputs: |x| <= sqrt(511) * gamma + beta
        z = rng.normal(0.0, 1.0, (rows, K))
        z[0, 0] = np.sqrt(511.0)                                  # the bound is attained
        h = np.maximum(z * gamma + beta, 0).astype(np.float32)
        w = rng.uniform(-w_mag, w_mag, (cols, K)).astype(np.float32)
        sw = _scale_exp(np.abs(w).max(), 14)
        sx = _scale_exp(22.63 * gamma + beta, 13)
        hs, ws = np.ldexp(h, sx).astype(np.float32), np.ldexp(w, sw).astype(np.float32)
        assert np.abs(ws).max() < 2.0 ** 15 and hs.max() < 2.0 ** 14 <= 65504          # inside fp16's range
        x1, x2 = _fp16_split(hs)
        w1, w2 = _fp16_split(ws)
        assert np.all(np.isfinite(x1.astype(np.float32))) and np.all(np.isfinite(w1.astype(np.float32)))
        f = lambda t: t.astype(np.float64)
        acc = f(x2) @ f(w1).T + f(x1) @ f(w2).T + f(x1) @ f(w1).T
        got = acc * 2.0 ** -(sw + sx)
        exact = f(h) @ f(w).T
        scale = np.abs(f(h)) @ np.abs(f(w)).T
        assert np.max(np.abs(got - exact) / scale) < 2.0 ** -20, (w_mag, gamma, beta)
    # a single FP16 pass is three orders of magnitude worse
    one = (f(x1) @ f(w1).T) * 2.0 ** -(sw + sx)
    assert np.max(np.abs(one - exact) / scale) > 2.0 ** -14
