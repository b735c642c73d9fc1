"""The per-frame checksum of include/drr.h (drr_read_checksums) in numpy -- pure Python, no native library.

The frame's little-endian u32 words (zero-padded) are taken in groups of 12 (48 bytes = 16 pixels):
    s_g = sum_j w[12 g + j] * ((2 j + 1) * C)  mod 2^32,      checksum = sum_g s_g * ((g + 1) * C mod 2^32)  mod 2^64,
with C = 0x9E3779B1.
"""
from __future__ import annotations

import numpy as np

C = 0x9E3779B1
GROUP = 12


def checksum_numpy(frame: np.ndarray) -> int:
    b = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    pad = (-b.size) % (4 * GROUP)
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view("<u4").astype(np.uint64).reshape(-1, GROUP)
    mj = (np.arange(GROUP, dtype=np.uint64) * np.uint64(2) + np.uint64(1)) * np.uint64(C) & np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        s = (w * mj).sum(axis=1, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
        k = (np.arange(1, s.size + 1, dtype=np.uint64) * np.uint64(C)) & np.uint64(0xFFFFFFFF)
        return int((s * k).sum(dtype=np.uint64))
