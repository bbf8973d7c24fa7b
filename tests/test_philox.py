"""Known-answer tests of the oracle's Philox4x32-10 (Random123 kat_vectors) and of the mask rule."""

import numpy as np

from oracle import philox


def _hex(words):
    return " ".join(f"{int(w):08x}" for w in words)


def test_random123_known_answers():
    z = np.uint32(0)
    assert _hex(philox.philox4x32_10(z, z, z, z, 0, 0)) == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    f = np.uint32(0xFFFFFFFF)
    assert _hex(philox.philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)) == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    c = [np.uint32(v) for v in (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)]
    assert _hex(philox.philox4x32_10(*c, 0xA4093822, 0x299F31D0)) == "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32, size=(4, 17), dtype=np.uint64).astype(np.uint32)
    vec = philox.philox4x32_10(c[0], c[1], c[2], c[3], 7, 9)
    for i in range(17):
        one = philox.philox4x32_10(c[0, i], c[1, i], c[2, i], c[3, i], 7, 9)
        assert [int(w[i]) for w in vec] == [int(w) for w in one]


def test_keep_rate_and_threshold():
    assert philox.keep_threshold(0.0) == 0
    assert philox.keep_threshold(0.5) == 2**15
    assert philox.keep_threshold(1.0) == 65535
    keep = philox.dropout_keep(42, 0.3, np.arange(64), np.zeros(64, dtype=np.uint32), 0, 0, 8, 64)
    assert keep.shape == (64, 8, 64)
    assert abs(keep.mean() - 0.7) < 0.01
    # independent of how sequences are batched
    part = philox.dropout_keep(42, 0.3, np.arange(10, 20), np.zeros(10, dtype=np.uint32), 0, 0, 8, 64)
    assert np.array_equal(part, keep[10:20])
