"""Presentation / export: the reference shows Pixels.pixels through an SDL texture (src/game.rs:521-525);
headless, the same RGB24 frame is written as a PNG.  Minimal encoder (zlib + CRC from the standard library)."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(rgb: np.ndarray, level: int = 6) -> bytes:
    """rgb: (H, W, 3) uint8 == Pixels.pixels reshaped.  Returns the PNG file contents (8-bit truecolour, filter 0)."""
    a = np.ascontiguousarray(rgb, np.uint8)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an (H, W, 3) uint8 array")
    h, w, _ = a.shape
    raw = np.zeros((h, 1 + w * 3), np.uint8)
    raw[:, 1:] = a.reshape(h, w * 3)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))


def decode_png(data: bytes) -> np.ndarray:
    """Inverse of encode_png (only what encode_png produces: 8-bit RGB, no interlace, filter type 0 on every row)."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG")
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] != (zlib.crc32(tag + body) & 0xFFFFFFFF):
            raise ValueError("bad chunk CRC")
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if (depth, ctype, interlace) != (8, 2, 0):
                raise ValueError("unsupported PNG flavour")
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 3)
    if raw[:, 0].any():
        raise ValueError("unsupported PNG filter")
    return raw[:, 1:].reshape(h, w, 3).copy()


def write_png(path: str, rgb: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(rgb))
