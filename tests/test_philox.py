"""Philox4x32-10: published known-answer vectors (Random123 kat_vectors) and host/library agreement."""
import numpy as np

from hygeia_b200 import philox


def test_known_answers():
    assert [int(x) for x in philox.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(x) for x in philox.philox4x32_10(f, f, f, f, f, f)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_range_and_determinism():
    u = philox.uniforms_by_site(12345, 3, 10000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(u[100:200], philox.uniforms_by_site(12345, 3, 100, t0=100))
    assert not np.array_equal(u[:100], philox.uniforms_by_site(12345, 4, 100))


def test_library_matches_numpy(built):
    from hygeia_b200 import _lib
    lib = _lib.load()
    u = philox.uniforms_by_site(987654321012345, 17, 64, t0=2**33)
    v = np.array([lib.hyg_philox_uniform(987654321012345, 17, 2**33 + t) for t in range(64)])
    assert np.array_equal(u, v)
