"""TEST INFRASTRUCTURE (oracle side) -- numpy restatement of the engine's noise source.

onnxruntime's ``RandomNormalLike`` stream cannot be reproduced (SURVEY.md §4, §7
hard part 9), so reference parity is defined at noise 0.  For noise > 0 the
engine uses a counter-based generator specified here, so that engine and oracle
can still be compared sample for sample:

    Philox4x32-10, key = (seed & 0xffffffff, seed >> 32),
    counter = (position, channel, stream, utterance_row)
    u1 = ((x0 >> 9) + 0.5) * 2**-23 ; u2 = ((x1 >> 9) + 0.5) * 2**-23
    n  = sqrt(-2 ln u1) * cos(2 pi u2)            (fp32)

stream 0: duration-predictor noise z (B,2,T); stream 1: prior noise (B,I,F).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        kk0 = np.uint64((k0 + r * W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ kk0, lo1, hi0 ^ c3 ^ kk1, lo0
    return c0, c1, c2, c3


def normal(seed: int, stream: int, row: int, position, channel) -> np.ndarray:
    """fp32 standard normals for broadcast(position, channel)."""
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    x0, x1, _, _ = philox4x32_10(position, channel, stream, row, k0, k1)
    u1 = ((x0 >> np.uint64(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    u2 = ((x1 >> np.uint64(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    return (r * np.cos(np.float32(6.283185307179586) * u2)).astype(np.float32)
