"""Benchmark / test workloads: BASELINE.json's configs made concrete, and where their content comes
from.  Content is the deterministic synthetic IWAD of synth_wad.py, or -- for the E1M1-class workloads -- a real IWAD when
the user sets DRR_WAD=<path to doom1.wad> (BASELINE.md:47; the format the reference reads: src/wad.rs:86-109).
"""
from __future__ import annotations

import math
import os
import struct
import tempfile

import numpy as np

from . import synth_wad

# name: (content, W, H, viewpoints per GPU, phase mask, passes of the batch per timed step, description)
WORKLOADS = {
    "walk320": ("e1m1", 320, 200, 4096, 3, 16, "BASELINE configs[1]: E1M1-class walk, 4096 viewpoints, 320x200, walls+flats+sky"),
    "walk1280": ("e1m1", 1280, 800, 512, 7, 12, "E1M1-class walk, 512 viewpoints, 1280x800, all phases (north_star target resolution)"),
    "walls1280": ("e1m1", 1280, 800, 256, 1, 24, "BASELINE configs[2]: walls only, 1280x800"),
    "flats1280": ("e1m1", 1280, 800, 256, 2, 24, "BASELINE configs[2]: flats+sky only, 1280x800"),
    "empty1280": ("e1m1", 1280, 800, 256, 0, 24, "no ops at all, 1280x800: clears + write-out only (fixed cost of a tile)"),
    "empty320": ("e1m1", 320, 200, 4096, 0, 16, "no ops at all, 320x200: clears + write-out only (fixed cost of a tile)"),
    "things640": ("e1m1", 640, 400, 4096, 7, 4, "BASELINE configs[3]: things, masked mids, lighting, 640x400, 4096 viewpoints per GPU"),
    "stress1920": ("stress", 1920, 1200, 8192, 7, 1, "BASELINE configs[4]: stress map at 1920x1200, 8192 viewpoints per GPU (65536 over 8)"),
}


def real_wad() -> str | None:
    """DRR_WAD=<path>: a real IWAD to use instead of the synthetic E1M1-class one."""
    p = os.environ.get("DRR_WAD")
    if p and not os.path.isfile(p):
        raise FileNotFoundError("DRR_WAD=%s: no such file" % p)
    return p or None


def wad_things(path: str, map_name: str = "E1M1") -> np.ndarray:
    """(x, y, angle in degrees, type, flags) of the map's THINGS lump (src/wad.rs:86-109 directory walk, src/map/things.rs)."""
    with open(path, "rb") as f:
        data = f.read()
    magic, n, ofs = struct.unpack_from("<4sii", data, 0)
    if magic not in (b"IWAD", b"PWAD"):
        raise ValueError("%s: not a WAD" % path)
    names = []
    for i in range(n):
        pos, size, name = struct.unpack_from("<ii8s", data, ofs + 16 * i)
        names.append((name.rstrip(b"\0").decode("ascii", "replace").upper(), pos, size))
    at = next((i for i, e in enumerate(names) if e[0] == map_name.upper()), None)
    if at is None:
        raise ValueError("%s: no map %s" % (path, map_name))
    for name, pos, size in names[at + 1:at + 12]:
        if name == "THINGS":
            return np.frombuffer(data, "<i2", size // 2, pos).reshape(-1, 5).astype(np.int32)
    raise ValueError("%s: %s has no THINGS lump" % (path, map_name))


def tour_viewpoints(things: np.ndarray, n: int) -> np.ndarray:
    """n viewpoints for a map we know nothing else about: the positions of its things (all inside the map), visited round
    after round, each round turned a little further (thing angle + round * golden angle)."""
    k = max(len(things), 1)
    out = np.zeros((n, 3), np.float32)
    for i in range(n):
        t = things[i % k]
        out[i] = (t[0], t[1], math.radians(float(t[2])) + (i // k) * 2.399963229728653)
    out[:, 2] = np.mod(out[:, 2] + math.pi, 2 * math.pi) - math.pi
    return out


class Content:
    """Where a workload's WAD is and how to get viewpoints for it."""

    def __init__(self, kind: str, cache_dir: str | None = None):
        self.kind = kind
        self.real = real_wad() if kind == "e1m1" else None
        if self.real:
            self.path, self.gm = self.real, None
            self.source = "real IWAD (DRR_WAD=%s)" % self.real
        else:
            data, self.gm, _ = synth_wad.build_wad(kind)
            d = cache_dir or tempfile.mkdtemp(prefix="drr_wad_")
            os.makedirs(d, exist_ok=True)
            self.path = os.path.join(d, "synth_%s.wad" % kind)
            if not os.path.exists(self.path) or open(self.path, "rb").read() != data:
                with open(self.path + ".tmp%d" % os.getpid(), "wb") as f:
                    f.write(data)
                os.replace(self.path + ".tmp%d" % os.getpid(), self.path)
            self.source = "synthetic IWAD (synth_wad.py, seed 0xD00D1993)"

    def viewpoints(self, n_total: int) -> np.ndarray:
        if self.real:
            return tour_viewpoints(wad_things(self.path), n_total)
        return synth_wad.walk_viewpoints(self.gm, n_total) if self.kind == "e1m1" else synth_wad.scatter_viewpoints(self.gm, n_total)
