"""Philox4x32-10 in numpy (test infrastructure; see oracle/__init__.py).

Restates the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
as 1, 2, 3", SC'11; Random123 `philox.h`): 10 rounds of two 32x32->64 multiplies with constants
0xD2511F53 / 0xCD9E8D57 and Weyl key increments 0x9E3779B9 / 0xBB67AE85.  Pinned by Random123's
known-answer vectors in tests/test_philox.py.  The CUDA twin is bayesformer_b200/csrc/common.cuh.
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 counters; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def keep_threshold(p: float) -> int:
    """16-bit decision threshold: keep iff half-word >= threshold; mirrors bfb::keep_threshold (common.cuh)."""
    t = float(np.float32(p)) * 65536.0
    if t <= 0.0:
        return 0
    if t >= 65535.0:
        return 65535
    return int(t)


def _keep8(words, thr: int) -> np.ndarray:
    """(..., 8) bool from the four uint32 output words of one call: bit j <- half (j & 1) of word (j >> 1)
    (bfb::keep8_of)."""
    thr = np.uint32(thr)
    halves = []
    for w in words:
        halves.append((w & np.uint32(0xFFFF)) >= thr)
        halves.append((w >> np.uint32(16)) >= thr)
    return np.stack(halves, axis=-1)


def dropout_keep(seed: int, p: float, items, draws, step: int, site: int, n_pos: int, d: int) -> np.ndarray:
    """Keep mask (n_seq, n_pos, d) bool for sequences given by parallel arrays (items, draws).

    Counter = (item, draw, (step << 16) | pos, (site << 16) | (feature >> 3)); feature f uses the 16-bit half
    f & 1 of word (f >> 1) & 3 (include/bayesformer_b200.h, `bfb200_masks`).
    """
    items = np.asarray(items, dtype=np.uint64)
    draws = np.asarray(draws, dtype=np.uint32)
    n_seq = items.shape[0]
    d8 = (d + 7) // 8
    c0 = (items & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None, None]
    c1 = draws[:, None, None]
    pos = np.arange(n_pos, dtype=np.uint32)[None, :, None]
    c2 = (np.uint32(step) << np.uint32(16)) | pos
    f8 = np.arange(d8, dtype=np.uint32)[None, None, :]
    c3 = (np.uint32(site) << np.uint32(16)) | f8
    words = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    keep = _keep8(words, keep_threshold(p)).reshape(n_seq, n_pos, d8 * 8)
    return keep[:, :, :d]


def sampling_uniform(seed: int, items, draws, step: int) -> np.ndarray:
    """U(0,1) float32 per sequence for the sampling scheme's token draw (site 0xFFFF, word 0, top 24 bits)."""
    items = np.asarray(items, dtype=np.uint64)
    draws = np.asarray(draws, dtype=np.uint32)
    c0 = (items & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    w = philox4x32_10(c0, draws, np.uint32(step << 16), np.uint32(0xFFFF << 16), seed & 0xFFFFFFFF,
                      (seed >> 32) & 0xFFFFFFFF)[0]
    return ((w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def mlp_keep(seed: int, p: float, items, n_draws: int, hidden: int) -> np.ndarray:
    """Keep mask (T, B, hidden) of the MLP hidden-layer dropout (site 0x100, step 0, pos 0)."""
    items = np.asarray(items, dtype=np.uint64)
    h8 = (hidden + 7) // 8
    c0 = (items & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None, :, None]
    c1 = np.arange(n_draws, dtype=np.uint32)[:, None, None]
    c3 = (np.uint32(0x100) << np.uint32(16)) | np.arange(h8, dtype=np.uint32)[None, None, :]
    words = philox4x32_10(c0, c1, np.uint32(0), c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    keep = _keep8(words, keep_threshold(p)).reshape(n_draws, items.shape[0], h8 * 8)
    return keep[:, :, :hidden]
