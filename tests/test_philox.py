"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) and the
distribution of the derived normals."""
import numpy as np

from oracle import philox


def _h(*a):
    return [int(x) for x in philox.philox4x32_10(*a)]


def test_random123_known_answers():
    assert _h(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _h(f, f, f, f, f, f) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _h(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_are_standard_and_counter_addressed():
    z = philox.normals(1870300, philox.KIND_ES, 0, 0, np.arange(32), 8192)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert np.isfinite(z).all()
    # any (member, index) is reproducible on its own
    z1 = philox.normals(1870300, philox.KIND_ES, 0, 0, [17], 8192)
    assert np.array_equal(z1[0], z[17])
    # different role / kind / generation decorrelate
    z2 = philox.normals(1870300, philox.KIND_GA, 0, 0, [17], 8192)
    z3 = philox.normals(1870300, philox.KIND_ES, 1, 0, [17], 8192)
    z4 = philox.normals(1870300, philox.KIND_ES, 0, 1, [17], 8192)
    for other in (z2, z3, z4):
        assert abs(np.corrcoef(z1[0], other[0])[0, 1]) < 0.05


def test_init_states_distribution():
    s = philox.init_states(1870300, 0, 20000)
    assert set(np.unique(s[:, 0])) == {0.0, 1.0}
    assert abs(s[:, 0].mean() - 0.5) < 0.02
    assert np.all(np.abs(s[:, 1:]) < 1) and abs(s[:, 1:].mean()) < 0.01
    assert abs(s[:, 1:].std() - 1 / np.sqrt(3)) < 0.01
