"""The per-frame checksum of include/drr.h (drr_read_checksums) in numpy -- pure Python, no native library.

sum over the frame's little-endian u32 words w_i of  w_i * ((i + 1) * 0x9E3779B1 mod 2^32),  mod 2^64.
"""
from __future__ import annotations

import numpy as np


def checksum_numpy(frame: np.ndarray) -> int:
    b = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    pad = (-b.size) % 4
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view("<u4").astype(np.uint64)
    k = (np.arange(1, w.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        return int((w * k).sum(dtype=np.uint64))
