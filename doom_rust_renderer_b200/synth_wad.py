"""Deterministic synthetic IWAD generator (host-side tooling; not on the render path).

There is no doom1.wad in this image or on the GPU box, and BASELINE.json allows "a synthetic WAD with the
same lump layout".  This module writes an IWAD that satisfies every assumption of the reference's loaders:

  * header / directory                         src/wad.rs:86-157  (magic must be "IWAD", S_START/S_END required)
  * map marker followed by 10 lumps in order   src/wad.rs:8-19,175-183
  * THINGS 10 B, LINEDEFS 14 B, SIDEDEFS 30 B, VERTEXES 4 B, SEGS 12 B, SSECTORS 4 B, NODES 28 B, SECTORS 26 B
                                               src/map/*.rs
  * PLAYPAL, PNAMES, TEXTURE1, picture-format patches and sprites, raw 4096-byte flats
                                               src/graphics/*.rs
  * sprite lump names NNNNFR[FR]               src/graphics/sprites.rs:38-56
  * sky texture SKY1, 256x128                  src/game.rs:199-227, src/renderer/visplanes.rs:49-50

Maps are built on a square grid, every sector boundary lies on a grid line, and the BSP is a k-d tree whose
partition lines are grid lines, so no seg is ever split.  Each maximal
horizontal run of same-sector cells in a grid row is one convex subsector and always owns at least two segs.

Everything is driven by a private PCG32 stream seeded with 0xD00D1993, so the bytes are reproducible.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

SEED = 0xD00D1993


class PCG32:
    """Minimal PCG-XSH-RR 64/32 (O'Neill).  Private so that the WAD bytes never depend on numpy/python RNG versions."""

    def __init__(self, seed: int, seq: int = 54):
        self.state = 0
        self.inc = ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self.next()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self.next()

    def next(self) -> int:
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def below(self, n: int) -> int:
        return self.next() % n if n > 0 else 0

    def rng(self, lo: int, hi: int) -> int:  # inclusive
        return lo + self.below(hi - lo + 1)

    def chance(self, p: float) -> bool:
        return self.next() < p * 4294967296.0

    def choice(self, seq):
        return seq[self.below(len(seq))]

    def bytes_array(self, n: int) -> np.ndarray:
        out = np.empty((n + 3) // 4, dtype=np.uint32)
        for i in range(out.size):
            out[i] = self.next()
        return out.view(np.uint8)[:n].copy()


def _name8(s: str) -> bytes:
    b = s.encode("ascii")
    assert len(b) <= 8, s
    return b + b"\0" * (8 - len(b))


# ----------------------------------------------------------------------------------------------------------
# graphics
# ----------------------------------------------------------------------------------------------------------
def encode_picture(img: np.ndarray, left_offset: int = 0, top_offset: int = 0) -> bytes:
    """img: int16 [h][w], -1 = transparent.  Doom picture format (doomwiki 'Picture format'; pictures.rs:100-126)."""
    h, w = img.shape
    assert h < 255
    cols = []
    for x in range(w):
        col = bytearray()
        y = 0
        while y < h:
            if img[y, x] < 0:
                y += 1
                continue
            y0 = y
            while y < h and img[y, x] >= 0 and y - y0 < 128:
                y += 1
            data = bytes(int(v) for v in img[y0:y, x])
            col += bytes([y0, y - y0, 0]) + data + b"\0"
        col.append(0xFF)
        cols.append(bytes(col))
    header = struct.pack("<hhhh", w, h, left_offset, top_offset)
    off = 8 + 4 * w
    table = bytearray()
    for c in cols:
        table += struct.pack("<I", off)
        off += len(c)
    return header + bytes(table) + b"".join(cols)


def _pattern(rng: PCG32, w: int, h: int, kind: int, base: int) -> np.ndarray:
    """A w x h palette-index pattern.  Deliberately high-frequency so that any texel-coordinate error shows up."""
    yy, xx = np.mgrid[0:h, 0:w]
    noise = rng.bytes_array(w * h).reshape(h, w).astype(np.int32)
    if kind == 0:  # bricks
        row = yy // 16
        xs = (xx + (row % 2) * 16) % 32
        v = base + (row * 7 + (xx + (row % 2) * 16) // 32 * 3) % 24 + (noise & 3)
        v = np.where((yy % 16 == 0) | (xs == 0), base + 40, v)
    elif kind == 1:  # vertical panels
        v = base + (xx // 8 * 5) % 32 + (yy // 32) * 2 + (noise & 1)
        v = np.where(xx % 32 == 31, base + 50, v)
    elif kind == 2:  # diagonal stripes
        v = base + ((xx + 2 * yy) // 4) % 48 + (noise & 1)
    elif kind == 3:  # checker + gradient
        v = base + ((xx // 8 + yy // 8) % 2) * 20 + yy // 4 % 16 + (noise & 3)
    else:  # pure noise band
        v = base + (noise % 64)
    return (v % 256).astype(np.int16)


def make_palette(rng: PCG32) -> bytes:
    # 256 seeded RGB triples; a few fixed entries so that saturation/zero cases are exercised
    pal = rng.bytes_array(768)
    pal[0:3] = (0, 0, 0)
    pal[3:6] = (255, 255, 255)
    pal[6:9] = (255, 0, 1)
    pal[9:12] = (1, 254, 128)
    return bytes(pal) * 14  # PLAYPAL holds 14 palettes; only palette 0 is read (palette.rs:11-28)


# thing types placed in synthetic maps: doomednum -> (sprite prefix, 8 rotations?, w, h)
THING_SPRITES = {
    3004: ("POSS", True, 41, 56),   # zombieman, rotations with mirrored lumps (A2A8 ..)
    3001: ("TROO", True, 43, 58),   # imp, eight separate lumps
    2035: ("BAR1", False, 23, 33),  # barrel
    2028: ("COLU", False, 23, 49),  # floor lamp, full-bright state
    2014: ("BON1", False, 15, 19),  # health bonus
    2012: ("MEDI", False, 28, 20),  # medikit
    35: ("CBRA", False, 29, 61),    # candelabra, full-bright state
    48: ("ELEC", False, 39, 127),   # tall techno pillar
}


def _sprite_image(rng: PCG32, w: int, h: int, base: int, variant: int) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w]
    cx = (w - 1) / 2.0
    # a blob whose outline depends on the variant (rotation) so that mirrored/rotated frames differ
    half = (w / 2.0) * (0.45 + 0.5 * np.abs(np.sin((yy + 3 * variant) * 0.23 + variant)))
    body = np.abs(xx - cx - (variant % 3 - 1) * 1.5) <= half
    noise = rng.bytes_array(w * h).reshape(h, w).astype(np.int32)
    v = base + (xx * 3 + yy * 5 + variant * 11) % 40 + (noise & 3)
    img = np.where(body, v % 256, -1).astype(np.int16)
    img[(yy % 9 == 4) & (xx % 7 == 3)] = -1  # pin-holes: several posts per column
    img[h - 1, :] = np.where(np.abs(xx[0] - cx) <= w / 4.0, base % 256, -1)
    return img


@dataclass
class TextureDef:
    name: str
    width: int
    height: int
    patches: list  # (origin_x, origin_y, patch_index)


def make_graphics(rng: PCG32):
    """Returns (patch lumps [(name, bytes)], texture defs, flats [(name, bytes)], sprite lumps [(name, bytes)])."""
    patches = []  # (name, img)

    def add_patch(img) -> int:
        patches.append(("PAT%03d" % len(patches), img))
        return len(patches) - 1

    textures: list[TextureDef] = []
    # 1) single-patch walls of assorted sizes (incl. non-power-of-two heights 72 / 96)
    sizes = [(64, 128), (128, 128), (64, 128), (128, 128), (256, 128), (64, 72), (128, 96), (64, 64),
             (32, 128), (16, 128), (64, 128), (128, 128), (64, 96), (128, 72), (64, 128), (128, 128)]
    for i, (w, h) in enumerate(sizes):
        p = add_patch(_pattern(rng, w, h, i % 5, (i * 37) % 200))
        textures.append(TextureDef("WALL%02d" % i, w, h, [(0, 0, p)]))
    # 2) multi-patch walls: side by side, overlapping, negative origins, patch running off the edge
    for i in range(12):
        w, h = rng.choice([(128, 128), (256, 128), (64, 128), (128, 96)])
        plist = []
        x = -rng.rng(0, 12)
        while x < w:
            pw = rng.choice([32, 64, 128])
            # patches are taller than the texture and start at or above its top edge: composite WALLS stay fully opaque, as
            # in real IWADs (holes only occur in the dedicated HOLEY00 / PARTIAL0 / GRATExx textures below)
            p = add_patch(_pattern(rng, pw, h + 16, rng.below(5), rng.below(200)))
            plist.append((x, -rng.rng(0, 16) if i % 3 == 0 else 0, p))
            x += pw - (rng.rng(0, 10) if i % 2 else 0)
        textures.append(TextureDef("COMP%02d" % i, w, h, plist))
    # 3) a wall whose LAST patch has transparent texels: they punch holes into the earlier patch (quirk Q1)
    base = add_patch(_pattern(rng, 128, 128, 0, 90))
    holes = _pattern(rng, 64, 64, 3, 10)
    yy, xx = np.mgrid[0:64, 0:64]
    holes[((xx // 8 + yy // 8) % 2) == 0] = -1
    hp = add_patch(holes)
    textures.append(TextureDef("HOLEY00", 128, 128, [(0, 0, base), (32, 32, hp)]))
    # 4) masked mid-textures (grates / fences): mostly transparent
    for i, (w, h) in enumerate([(64, 128), (128, 128), (64, 72), (32, 128)]):
        img = _pattern(rng, w, h, 1, 60 + 30 * i)
        yy, xx = np.mgrid[0:h, 0:w]
        img[((xx % 16) >= 4) & ((yy % 16) >= 4)] = -1
        if i == 2:
            img[:, 0:5] = -1  # fully transparent columns
        p = add_patch(img)
        textures.append(TextureDef("GRATE%02d" % i, w, h, [(0, 0, p)]))
    # 5) a texture with uncovered (None) area: patch smaller than the texture
    p = add_patch(_pattern(rng, 64, 64, 2, 120))
    textures.append(TextureDef("PARTIAL0", 128, 128, [(16, 16, p), (70, 50, p)]))
    # 6) fillers up to 40 wall textures + the sky
    while len(textures) < 40:
        i = len(textures)
        w, h = rng.choice([(64, 128), (128, 128), (256, 128)])
        p = add_patch(_pattern(rng, w, h, i % 5, (i * 53) % 220))
        textures.append(TextureDef("FILL%02d" % i, w, h, [(0, 0, p)]))
    sky = np.zeros((128, 256), dtype=np.int16)
    yy, xx = np.mgrid[0:128, 0:256]
    sky[:, :] = (160 + ((xx // 4 + yy // 2) % 64) + (rng.bytes_array(128 * 256).reshape(128, 256) & 1)) % 256
    sp = add_patch(sky[:, :128].copy())
    sp2 = add_patch(sky[:, 128:].copy())
    textures.append(TextureDef("SKY1", 256, 128, [(0, 0, sp), (128, 0, sp2)]))

    patch_lumps = [(n, encode_picture(img)) for n, img in patches]

    flats = []
    flat_names = ["FLOOR%02d" % i for i in range(12)] + ["CEIL%02d" % i for i in range(8)] + ["NUKAGE1", "NUKAGE2", "NUKAGE3", "F_SKY1"]
    for i, n in enumerate(flat_names):
        img = _pattern(rng, 64, 64, i % 5, (i * 29) % 230).astype(np.uint8)
        flats.append((n, img.tobytes()))

    sprites = []
    for num, (prefix, rotated, w, h) in sorted(THING_SPRITES.items()):
        base = (num * 13) % 200
        lo, to = w // 2, h - 4
        if not rotated:
            sprites.append((prefix + "A0", encode_picture(_sprite_image(rng, w, h, base, 0), lo, to)))
        elif prefix == "POSS":
            for nm, var in (("A1", 1), ("A2A8", 2), ("A3A7", 3), ("A4A6", 4), ("A5", 5)):
                sprites.append((prefix + nm, encode_picture(_sprite_image(rng, w, h, base, var), lo, to)))
        else:
            for r in range(1, 9):
                sprites.append((prefix + "A%d" % r, encode_picture(_sprite_image(rng, w + (r % 3), h - (r % 2), base, r), lo, to)))
    return patch_lumps, textures, flats, sprites


def encode_textures(textures: list[TextureDef], n_patches: int) -> bytes:
    """TEXTURE1 lump (textures.rs:208-255): count, offsets, then maptexture_t records with 10-byte patches."""
    recs = []
    for t in textures:
        r = _name8(t.name) + struct.pack("<IhhIh", 0, t.width, t.height, 0, len(t.patches))
        for (ox, oy, p) in t.patches:
            assert 0 <= p < n_patches
            r += struct.pack("<hhhhh", ox, oy, p, 1, 0)
        recs.append(r)
    off = 4 + 4 * len(recs)
    head = struct.pack("<I", len(recs))
    for r in recs:
        head += struct.pack("<I", off)
        off += len(r)
    return head + b"".join(recs)


# ----------------------------------------------------------------------------------------------------------
# map
# ----------------------------------------------------------------------------------------------------------
@dataclass
class SectorDef:
    floor: int
    ceil: int
    floor_flat: str
    ceil_flat: str
    light: int
    wall_tex: str = "WALL00"
    special: int = 0
    step_tex: str = "WALL01"
    mid_tex: str = "-"  # masked mid texture put on two-sided lines whose FRONT is this sector


@dataclass
class GridMap:
    gw: int
    gh: int
    cell: int
    x0: int
    y0: int
    sectors: list = field(default_factory=list)
    cells: np.ndarray = None  # int32 [gh][gw], -1 = void
    things: list = field(default_factory=list)  # (x, y, angle_deg, type, flags)
    rooms: list = field(default_factory=list)  # (cx0, cy0, cx1, cy1, sector) for path scripting

    def wx(self, cx):
        return self.x0 + cx * self.cell

    def wy(self, cy):
        return self.y0 + cy * self.cell


def _carve(gm: GridMap, cx0, cy0, cx1, cy1, sector: int):
    gm.cells[cy0:cy1, cx0:cx1] = sector


# opaque wall textures used at random; the two holey ones (HOLEY00, PARTIAL0) are assigned to a few specific sectors below,
# like the rare see-through walls of real maps, so that they are exercised without dominating the workload
WALL_POOL = ["WALL%02d" % i for i in range(16)] + ["COMP%02d" % i for i in range(12)] + ["FILL%02d" % i for i in range(35, 40)]
FLOOR_POOL = ["FLOOR%02d" % i for i in range(12)]
CEIL_POOL = ["CEIL%02d" % i for i in range(8)]
GRATES = ["GRATE00", "GRATE01", "GRATE02", "GRATE03"]


def build_e1m1_class(rng: PCG32) -> GridMap:
    """An E1M1-*class* map: rooms joined by corridors and stairs, an outdoor sky courtyard with a sky-hack border,
    platforms, pits with animated nukage, light panels, pillars, windows with masked grates, one closed door."""
    gm = GridMap(gw=60, gh=48, cell=64, x0=-1920, y0=-1536)
    gm.cells = -np.ones((gm.gh, gm.gw), dtype=np.int32)

    def new_sector(**kw) -> int:
        gm.sectors.append(SectorDef(**kw))
        return len(gm.sectors) - 1

    def room_sector(sky=False, tall=False) -> int:
        floor = 8 * rng.rng(-4, 6)
        height = 8 * (rng.rng(20, 32) if tall else rng.rng(12, 20))
        return new_sector(floor=floor, ceil=floor + height, floor_flat=rng.choice(FLOOR_POOL),
                          ceil_flat="F_SKY1" if sky else rng.choice(CEIL_POOL), light=rng.rng(96, 255),
                          wall_tex=rng.choice(WALL_POOL), step_tex=rng.choice(WALL_POOL))

    # --- rooms on a jittered 5 x 4 lattice -------------------------------------------------------------
    rooms = []
    for ry in range(4):
        for rx in range(5):
            w, h = rng.rng(5, 9), rng.rng(4, 8)
            cx0 = 1 + rx * 12 + rng.rng(0, 12 - w - 1)
            cy0 = 1 + ry * 12 + rng.rng(0, 12 - h - 1)
            sky = (rx, ry) == (2, 1)
            if sky:
                w, h, cx0, cy0 = 10, 9, 1 + rx * 12, 1 + ry * 12 + 1
            s = room_sector(sky=sky, tall=sky)
            _carve(gm, cx0, cy0, cx0 + w, cy0 + h, s)
            rooms.append((cx0, cy0, cx0 + w, cy0 + h, s))
    gm.rooms = rooms

    gm.sectors[rooms[7][4]].wall_tex = "HOLEY00"
    gm.sectors[rooms[13][4]].wall_tex = "PARTIAL0"
    gm.sectors[rooms[18][4]].step_tex = "HOLEY00"

    # --- corridors: lattice neighbours, L-shaped, some as staircases -------------------------------------
    def corridor(a, b, stairs: bool):
        ax, ay = (a[0] + a[2]) // 2, (a[1] + a[3]) // 2
        bx, by = (b[0] + b[2]) // 2, (b[1] + b[3]) // 2
        path = []
        x, y = ax, ay
        while x != bx:
            path.append((x, y))
            x += 1 if bx > x else -1
        while y != by:
            path.append((x, y))
            y += 1 if by > y else -1
        path.append((bx, by))
        path = [(x, y) for (x, y) in path if gm.cells[y, x] < 0]
        if not path:
            return
        sa, sb = gm.sectors[a[4]], gm.sectors[b[4]]
        if stairs and len(path) >= 3:
            n = len(path)
            tex = rng.choice(WALL_POOL)
            for i, (x, y) in enumerate(path):
                f = sa.floor + (sb.floor - sa.floor) * (i + 1) // (n + 1)
                f = (f // 4) * 4
                s = new_sector(floor=f, ceil=max(sa.ceil, sb.ceil), floor_flat=sa.floor_flat, ceil_flat=rng.choice(CEIL_POOL),
                               light=(sa.light * (n - i) + sb.light * i) // n, wall_tex=tex, step_tex=rng.choice(WALL_POOL))
                gm.cells[y, x] = s
        else:
            f = min(sa.floor, sb.floor)
            s = new_sector(floor=f, ceil=f + 8 * rng.rng(9, 13), floor_flat=rng.choice(FLOOR_POOL), ceil_flat=rng.choice(CEIL_POOL),
                           light=rng.rng(96, 200), wall_tex=rng.choice(WALL_POOL), step_tex=rng.choice(WALL_POOL))
            for (x, y) in path:
                gm.cells[y, x] = s
                if rng.chance(0.35) and x + 1 < gm.gw and gm.cells[y, x + 1] < 0 and gm.cells[y, min(x + 2, gm.gw - 1)] < 0:
                    gm.cells[y, x + 1] = s  # widen

    def R(rx, ry):
        return rooms[ry * 5 + rx]

    for ry in range(4):
        for rx in range(5):
            if rx + 1 < 5:
                corridor(R(rx, ry), R(rx + 1, ry), stairs=rng.chance(0.3))
            if ry + 1 < 4 and (rx % 2 == 0 or rng.chance(0.5)):
                corridor(R(rx, ry), R(rx, ry + 1), stairs=rng.chance(0.4))

    # --- interior features --------------------------------------------------------------------------------
    for (cx0, cy0, cx1, cy1, s) in rooms:
        sd = gm.sectors[s]
        w, h = cx1 - cx0, cy1 - cy0
        sky = sd.ceil_flat == "F_SKY1"
        if sky:
            # sky-hack border: an inner sky sector with a LOWER sky ceiling (segs.rs:463-477) and a raised planter
            inner = new_sector(floor=sd.floor + 16, ceil=sd.ceil - 64, floor_flat="FLOOR03", ceil_flat="F_SKY1", light=sd.light,
                               wall_tex=sd.wall_tex, step_tex="COMP03")
            _carve(gm, cx0 + 2, cy0 + 2, cx1 - 2, cy1 - 2, inner)
            pit = new_sector(floor=sd.floor - 24, ceil=sd.ceil - 64, floor_flat="NUKAGE1", ceil_flat="F_SKY1", light=sd.light,
                             wall_tex=sd.wall_tex, step_tex="WALL05")
            _carve(gm, cx0 + 4, cy0 + 4, cx1 - 4, cy1 - 4, pit)
            continue
        k = rng.below(6)
        if k == 0 and w >= 5 and h >= 4:  # raised platform
            p = new_sector(floor=sd.floor + 8 * rng.rng(1, 4), ceil=sd.ceil, floor_flat=rng.choice(FLOOR_POOL), ceil_flat=sd.ceil_flat,
                           light=min(255, sd.light + 32), wall_tex=sd.wall_tex, step_tex=rng.choice(WALL_POOL))
            _carve(gm, cx0 + 1, cy0 + 1, cx1 - 2, cy1 - 1, p)
        elif k == 1 and w >= 5 and h >= 5:  # nukage pit (animated flat, flats.rs:103-111)
            p = new_sector(floor=sd.floor - 16, ceil=sd.ceil, floor_flat=rng.choice(["NUKAGE1", "NUKAGE2", "NUKAGE3"]),
                           ceil_flat=sd.ceil_flat, light=sd.light, wall_tex=sd.wall_tex, step_tex=rng.choice(WALL_POOL))
            _carve(gm, cx0 + 2, cy0 + 2, cx1 - 2, cy1 - 2, p)
        elif k == 2 and w >= 4 and h >= 4:  # ceiling light panel: lower ceiling, bright
            p = new_sector(floor=sd.floor, ceil=sd.ceil - 16, floor_flat=sd.floor_flat, ceil_flat=rng.choice(CEIL_POOL), light=255,
                           wall_tex=sd.wall_tex, step_tex=rng.choice(WALL_POOL))
            _carve(gm, cx0 + 1, cy0 + 1, cx0 + 3, cy0 + 3, p)
        elif k == 3 and w >= 5 and h >= 5:  # pillars (void cells)
            gm.cells[cy0 + 1, cx0 + 1] = -1
            gm.cells[cy1 - 2, cx1 - 2] = -1
            if rng.chance(0.5):
                gm.cells[cy0 + 1, cx1 - 2] = -1
        elif k == 4 and w >= 6:  # fence: thin raised strip with a masked grate on it
            p = new_sector(floor=sd.floor + 8, ceil=sd.ceil - 8, floor_flat=sd.floor_flat, ceil_flat=sd.ceil_flat, light=sd.light,
                           wall_tex=sd.wall_tex, step_tex=rng.choice(WALL_POOL), mid_tex=rng.choice(GRATES))
            _carve(gm, cx0 + w // 2, cy0, cx0 + w // 2 + 1, cy1, p)
        # k == 5: plain room

    # --- windows / doors between horizontally adjacent rooms separated by exactly one void cell -------------
    n_doors = 0
    for y in range(1, gm.gh - 1):
        for x in range(1, gm.gw - 1):
            if gm.cells[y, x] >= 0:
                continue
            l, r = gm.cells[y, x - 1], gm.cells[y, x + 1]
            if l >= 0 and r >= 0 and l != r and gm.cells[y - 1, x] < 0 and gm.cells[y + 1, x] < 0 and rng.chance(0.5):
                a, b = gm.sectors[l], gm.sectors[r]
                lo, hi = max(a.floor, b.floor), min(a.ceil, b.ceil)
                if hi - lo < 48:
                    continue
                if n_doors == 0:  # one closed door: zero-height sector (segs.rs:222-225)
                    s = new_sector(floor=lo, ceil=lo, floor_flat=a.floor_flat, ceil_flat=a.ceil_flat, light=a.light, wall_tex="WALL09",
                                   step_tex="WALL10")
                    n_doors += 1
                else:
                    s = new_sector(floor=lo + 24, ceil=hi - 16, floor_flat=rng.choice(FLOOR_POOL), ceil_flat=rng.choice(CEIL_POOL),
                                   light=rng.rng(128, 255), wall_tex=rng.choice(WALL_POOL), step_tex=rng.choice(WALL_POOL),
                                   mid_tex=rng.choice(GRATES) if rng.chance(0.6) else "-")
                gm.cells[y, x] = s

    # --- things -------------------------------------------------------------------------------------------
    px, py = (rooms[0][0] + rooms[0][2]) // 2, (rooms[0][1] + rooms[0][3]) // 2
    while gm.cells[py, px] < 0:
        px += 1
    gm.things.append((gm.wx(px) + gm.cell // 2, gm.wy(py) + gm.cell // 2, 45, 1, 7))
    types = sorted(THING_SPRITES)
    tries = 0
    while len(gm.things) < 131 and tries < 5000:
        tries += 1
        cx, cy = rng.below(gm.gw), rng.below(gm.gh)
        if gm.cells[cy, cx] < 0:
            continue
        t = rng.choice(types)
        gm.things.append((gm.wx(cx) + rng.rng(12, gm.cell - 12), gm.wy(cy) + rng.rng(12, gm.cell - 12), 45 * rng.below(8), t, 7))
    return gm


def build_stress(rng: PCG32, n: int = 48, cell: int = 256) -> GridMap:
    """Config 5: n x n grid of `cell`-unit cells, random floor/ceiling steps (many visplanes), 512-tall rooms, sparse pillars."""
    gm = GridMap(gw=n + 2, gh=n + 2, cell=cell, x0=-(n + 2) * cell // 2, y0=-(n + 2) * cell // 2)
    gm.cells = -np.ones((gm.gh, gm.gw), dtype=np.int32)
    for cy in range(1, n + 1):
        for cx in range(1, n + 1):
            if rng.chance(0.06):
                continue  # pillar
            floor = 8 * rng.rng(0, 6)
            gm.sectors.append(SectorDef(floor=floor, ceil=512 - 8 * rng.rng(0, 6), floor_flat=rng.choice(FLOOR_POOL),
                                        ceil_flat="F_SKY1" if rng.chance(0.1) else rng.choice(CEIL_POOL), light=rng.rng(96, 255),
                                        wall_tex=rng.choice(WALL_POOL), step_tex=rng.choice(WALL_POOL)))
            gm.cells[cy, cx] = len(gm.sectors) - 1
    c = n // 2
    while gm.cells[c, c] < 0:
        c += 1
    gm.things.append((gm.wx(c) + cell // 2, gm.wy(c) + cell // 2, 90, 1, 7))
    types = sorted(THING_SPRITES)
    for _ in range(200):
        cx, cy = rng.rng(1, n), rng.rng(1, n)
        if gm.cells[cy, cx] >= 0:
            gm.things.append((gm.wx(cx) + rng.rng(16, cell - 16), gm.wy(cy) + rng.rng(16, cell - 16), 45 * rng.below(8), rng.choice(types), 7))
    gm.rooms = [(1, 1, n + 1, n + 1, 0)]
    return gm


def compile_map(gm: GridMap, rng: PCG32) -> dict:
    """Grid -> Doom map lumps with a grid-aligned k-d BSP.  Returns {lump name: bytes}."""
    cells, S = gm.cells, gm.cell
    gh, gw = cells.shape

    def sec(cx, cy):
        if cx < 0 or cy < 0 or cx >= gw or cy >= gh:
            return -1
        return int(cells[cy, cx])

    verts: dict = {}
    vlist = []

    def vid(x, y):
        k = (x, y)
        if k not in verts:
            verts[k] = len(vlist)
            vlist.append(k)
        return verts[k]

    # pieces: maximal horizontal runs per row
    piece_of = -np.ones((gh, gw), dtype=np.int32)
    pieces = []  # (row, cx0, cx1, sector)
    for cy in range(gh):
        cx = 0
        while cx < gw:
            s = sec(cx, cy)
            if s < 0:
                cx += 1
                continue
            c0 = cx
            while cx < gw and sec(cx, cy) == s:
                cx += 1
            piece_of[cy, c0:cx] = len(pieces)
            pieces.append((cy, c0, cx, s))
    piece_segs = [[] for _ in pieces]

    sidedefs, linedefs = [], []

    def add_sidedef(sector, upper, lower, middle, xo=0, yo=0):
        sidedefs.append((xo, yo, upper, lower, middle, sector))
        return len(sidedefs) - 1

    def add_line(p_start, p_end, front, back, cells_front, cells_back):
        """p_*: world points; front/back: sector ids (back may be -1); cells_*: list of (piece id, seg start pt, seg end pt)
        for unit stretches ordered from the linedef start to its end, on each side."""
        fs = gm.sectors[front]
        two = back >= 0
        flags = 4 if two else 1
        if rng.chance(0.3):
            flags |= 8
        if rng.chance(0.3):
            flags |= 16
        xo = rng.rng(-20, 40) if rng.chance(0.4) else 0
        yo = rng.rng(-16, 16) if rng.chance(0.3) else 0
        if two:
            bs = gm.sectors[back]
            mid_f = fs.mid_tex if fs.mid_tex != "-" else (bs.mid_tex if bs.mid_tex != "-" else "-")
            f_sd = add_sidedef(front, fs.step_tex, fs.step_tex, mid_f, xo, yo)
            b_sd = add_sidedef(back, bs.step_tex, bs.step_tex, mid_f if rng.chance(0.5) else "-", 0, 0)
        else:
            f_sd = add_sidedef(front, "-", "-", fs.wall_tex, xo, yo)
            b_sd = -1
        ld = len(linedefs)
        linedefs.append((vid(*p_start), vid(*p_end), flags, 0, 0, f_sd, b_sd))
        length_acc = 0
        for (pc, a, b) in cells_front:  # direction 0: offset measured from the linedef start
            piece_segs[pc].append((vid(*a), vid(*b), ld, 0, length_acc))
            length_acc += abs(b[0] - a[0]) + abs(b[1] - a[1])
        if two:
            length_acc = 0
            for (pc, a, b) in reversed(cells_back):  # direction 1: runs from the linedef end; offset from the linedef end
                piece_segs[pc].append((vid(*b), vid(*a), ld, 1, length_acc))
                length_acc += abs(b[0] - a[0]) + abs(b[1] - a[1])

    def merge_runs(stretches):
        """stretches: list of (piece id, a, b) unit pieces in order; merge consecutive ones with the same piece id."""
        out = []
        for (pc, a, b) in stretches:
            if out and out[-1][0] == pc and out[-1][2] == a:
                out[-1] = (pc, out[-1][1], b)
            else:
                out.append((pc, a, b))
        return out

    # horizontal edges: between row cy-1 (south) and cy (north), at y = wy(cy)
    for cy in range(gh + 1):
        cx = 0
        while cx < gw:
            n_s, s_s = sec(cx, cy), sec(cx, cy - 1)
            if n_s == s_s:
                cx += 1
                continue
            c0 = cx
            while cx < gw and sec(cx, cy) == n_s and sec(cx, cy - 1) == s_s:
                cx += 1
            y = gm.wy(cy)
            xa, xb = gm.wx(c0), gm.wx(cx)
            # choose the front side: the lower sector id if both exist (deterministic), else the existing one
            front_is_north = (n_s >= 0) and (s_s < 0 or n_s < s_s)
            if front_is_north:  # north on the right => direction west: start = east end
                unit = [(c, (gm.wx(c + 1), y), (gm.wx(c), y)) for c in range(cx - 1, c0 - 1, -1)]
                f_st = merge_runs([(int(piece_of[cy, c]), a, b) for (c, a, b) in unit])
                b_st = merge_runs([(int(piece_of[cy - 1, c]), a, b) for (c, a, b) in unit]) if s_s >= 0 else []
                add_line((xb, y), (xa, y), n_s, s_s, f_st, b_st)
            else:  # south on the right => direction east
                unit = [(c, (gm.wx(c), y), (gm.wx(c + 1), y)) for c in range(c0, cx)]
                f_st = merge_runs([(int(piece_of[cy - 1, c]), a, b) for (c, a, b) in unit])
                b_st = merge_runs([(int(piece_of[cy, c]), a, b) for (c, a, b) in unit]) if n_s >= 0 else []
                add_line((xa, y), (xb, y), s_s, n_s, f_st, b_st)
    # vertical edges: between column cx-1 (west) and cx (east), at x = wx(cx)
    for cx in range(gw + 1):
        cy = 0
        while cy < gh:
            e_s, w_s = sec(cx, cy), sec(cx - 1, cy)
            if e_s == w_s:
                cy += 1
                continue
            r0 = cy
            while cy < gh and sec(cx, cy) == e_s and sec(cx - 1, cy) == w_s and cy - r0 < 4:
                cy += 1
            x = gm.wx(cx)
            ya, yb = gm.wy(r0), gm.wy(cy)
            front_is_east = (e_s >= 0) and (w_s < 0 or e_s < w_s)
            if front_is_east:  # east on the right => direction north
                unit = [(r, (x, gm.wy(r)), (x, gm.wy(r + 1))) for r in range(r0, cy)]
                f_st = [(int(piece_of[r, cx]), a, b) for (r, a, b) in unit]
                b_st = [(int(piece_of[r, cx - 1]), a, b) for (r, a, b) in unit] if w_s >= 0 else []
                add_line((x, ya), (x, yb), e_s, w_s, f_st, b_st)
            else:  # west on the right => direction south
                unit = [(r, (x, gm.wy(r + 1)), (x, gm.wy(r))) for r in range(cy - 1, r0 - 1, -1)]
                f_st = [(int(piece_of[r, cx - 1]), a, b) for (r, a, b) in unit]
                b_st = [(int(piece_of[r, cx]), a, b) for (r, a, b) in unit] if e_s >= 0 else []
                add_line((x, yb), (x, ya), w_s, e_s, f_st, b_st)

    # SEGS / SSECTORS
    segs, ssectors = [], []
    for pi, lst in enumerate(piece_segs):
        assert len(lst) >= 2, "every piece owns its two end walls"
        ssectors.append((len(lst), len(segs)))
        for (a, b, ld, direction, off) in lst:
            ax, ay = vlist[a]
            bx, by = vlist[b]
            ang = int(round(math.atan2(by - ay, bx - ax) * 32768.0 / math.pi)) & 0xFFFF
            if ang >= 32768:
                ang -= 65536
            segs.append((a, b, ang, ld, direction, off))

    # NODES: k-d tree over rows, then over pieces within a row
    nodes = []
    rows = {}
    for pi, (cy, c0, c1, s) in enumerate(pieces):
        rows.setdefault(cy, []).append(pi)

    def bbox_of(plist):
        xs0 = min(gm.wx(pieces[p][1]) for p in plist)
        xs1 = max(gm.wx(pieces[p][2]) for p in plist)
        ys0 = min(gm.wy(pieces[p][0]) for p in plist)
        ys1 = max(gm.wy(pieces[p][0] + 1) for p in plist)
        return (ys1, ys0, xs0, xs1)  # top, bottom, left, right

    def build_row(plist):
        if len(plist) == 1:
            return 0x8000 | plist[0], plist
        m = len(plist) // 2
        west, east = plist[:m], plist[m:]
        lw, pw = build_row(west)
        le, pe = build_row(east)
        x = gm.wx(pieces[east[0]][1])
        y = gm.wy(pieces[east[0]][0])
        # partition pointing north: left = west, right = east  (vertexes.rs:32-34: cross <= 0 is "left")
        nodes.append((x, y, 0, S, bbox_of(pe), bbox_of(pw), le, lw))
        return len(nodes) - 1, plist

    def build_rows(rlist):
        if len(rlist) == 1:
            return build_row(rows[rlist[0]])
        m = len(rlist) // 2
        south, north = rlist[:m], rlist[m:]
        ls, ps = build_rows(south)
        ln, pn = build_rows(north)
        y = gm.wy(north[0])
        # partition pointing east: left = north, right = south
        nodes.append((gm.x0, y, S, 0, bbox_of(ps), bbox_of(pn), ls, ln))
        return len(nodes) - 1, ps + pn

    root, _ = build_rows(sorted(rows))
    if root & 0x8000:  # a single subsector: the reference needs at least one node (mod.rs:119)
        p = pieces[root & 0x7FFF]
        nodes.append((gm.wx(p[1]), gm.wy(p[0]), S, 0, bbox_of([root & 0x7FFF]), bbox_of([root & 0x7FFF]), root, root))
    assert len(vlist) < 32768 and len(segs) < 32768 and len(ssectors) < 32768 and len(nodes) < 32768

    def s16(v):
        return v - 65536 if v >= 32768 else v

    lumps = {}
    lumps["THINGS"] = b"".join(struct.pack("<hhhhh", *t) for t in gm.things)
    lumps["LINEDEFS"] = b"".join(struct.pack("<hhhhhhh", *l) for l in linedefs)
    lumps["SIDEDEFS"] = b"".join(struct.pack("<hh", xo, yo) + _name8(u) + _name8(lo) + _name8(m) + struct.pack("<h", s)
                                 for (xo, yo, u, lo, m, s) in sidedefs)
    lumps["VERTEXES"] = b"".join(struct.pack("<hh", x, y) for (x, y) in vlist)
    lumps["SEGS"] = b"".join(struct.pack("<hhhhhh", a, b, ang, ld, d, off) for (a, b, ang, ld, d, off) in segs)
    lumps["SSECTORS"] = b"".join(struct.pack("<hh", n, f) for (n, f) in ssectors)
    lumps["NODES"] = b"".join(struct.pack("<hhhh", x, y, dx, dy) + struct.pack("<hhhh", *rb) + struct.pack("<hhhh", *lb) +
                              struct.pack("<hh", s16(rc), s16(lc)) for (x, y, dx, dy, rb, lb, rc, lc) in nodes)
    lumps["SECTORS"] = b"".join(struct.pack("<hh", s.floor, s.ceil) + _name8(s.floor_flat) + _name8(s.ceil_flat) +
                                struct.pack("<hhh", s.light, s.special, 0) for s in gm.sectors)
    lumps["REJECT"] = b""
    lumps["BLOCKMAP"] = b""
    lumps["_stats"] = dict(sectors=len(gm.sectors), linedefs=len(linedefs), sidedefs=len(sidedefs), vertexes=len(vlist), segs=len(segs),
                           ssectors=len(ssectors), nodes=len(nodes), things=len(gm.things))
    return lumps


MAP_LUMP_ORDER = ["THINGS", "LINEDEFS", "SIDEDEFS", "VERTEXES", "SEGS", "SSECTORS", "NODES", "SECTORS", "REJECT", "BLOCKMAP"]


def build_wad(kind: str = "e1m1", seed: int = SEED):
    """Returns (wad bytes, GridMap, stats).  kind: 'e1m1' (E1M1-class) or 'stress' (config 5)."""
    rng = PCG32(seed)
    palette = make_palette(rng)
    patch_lumps, textures, flats, sprites = make_graphics(rng)
    if kind in ("e1m1", "e1m1_time"):
        gm = build_e1m1_class(rng)
    elif kind == "stress":
        gm = build_stress(rng)
    elif kind == "tiny":
        gm = build_stress(rng, n=6, cell=128)
    else:
        raise ValueError(kind)
    if kind == "e1m1_time":
        # the same map for the time axis: every third sector gets a light effect (flash, strobes, glow, synchronised
        # strobes, fire flicker: thinkers.rs:14-76) and the animated things get the sprite frames their states cycle through
        # (own generator stream, so that the base content stays what 'e1m1' has)
        for k, sec in enumerate(gm.sectors):
            if k % 3 == 1:
                sec.special = (1, 2, 3, 8, 12, 13, 17, 4)[(k // 3) % 8]
        rng2 = PCG32(seed + 1)
        for num, (prefix, rotated, w, h) in sorted(THING_SPRITES.items()):
            base, lo, to = (num * 13) % 200, w // 2, h - 4
            frames = {"BAR1": "B", "BON1": "BCD", "POSS": "B", "TROO": "B"}.get(prefix, "")
            for fi, fr in enumerate(frames):
                if not rotated:
                    sprites.append((prefix + fr + "0", encode_picture(_sprite_image(rng2, w, h, base + 17 * (fi + 1), 0), lo, to)))
                elif prefix == "POSS":
                    for nm, var in (("1", 1), ("2%s8" % fr, 2), ("3%s7" % fr, 3), ("4%s6" % fr, 4), ("5", 5)):
                        sprites.append((prefix + fr + nm, encode_picture(_sprite_image(rng2, w, h, base + 17, var), lo, to)))
                else:
                    for r in range(1, 9):
                        sprites.append((prefix + fr + "%d" % r, encode_picture(_sprite_image(rng2, w + (r % 3), h - (r % 2), base + 17, r), lo, to)))
    m = compile_map(gm, rng)
    stats = m.pop("_stats")

    lumps = [("PLAYPAL", palette), ("E1M1", b"")]
    lumps += [(n, m[n]) for n in MAP_LUMP_ORDER]
    lumps.append(("TEXTURE1", encode_textures(textures, len(patch_lumps))))
    lumps.append(("PNAMES", struct.pack("<I", len(patch_lumps)) + b"".join(_name8(n) for n, _ in patch_lumps)))
    lumps.append(("P_START", b""))
    lumps += patch_lumps
    lumps.append(("P_END", b""))
    lumps.append(("F_START", b""))
    lumps += flats
    lumps.append(("F_END", b""))
    lumps.append(("S_START", b""))
    lumps += sprites
    lumps.append(("S_END", b""))

    body = bytearray()
    directory = bytearray()
    off = 12
    for name, data in lumps:
        directory += struct.pack("<II", off if data else 0, len(data)) + _name8(name)
        body += data
        off += len(data)
    wad = b"IWAD" + struct.pack("<II", len(lumps), 12 + len(body)) + bytes(body) + bytes(directory)
    return wad, gm, stats


# ----------------------------------------------------------------------------------------------------------
# viewpoints
# ----------------------------------------------------------------------------------------------------------
def walk_viewpoints(gm: GridMap, n: int) -> np.ndarray:
    """n arc-length-uniform samples of a polyline through the room centres (over walkable cells),
    angle = path heading + 0.6*sin(2*pi*i/257).  Returns float32 [n][3] = (x, y, angle)."""
    from collections import deque
    cells = gm.cells
    gh, gw = cells.shape

    def centre(r):
        cx, cy = (r[0] + r[2]) // 2, (r[1] + r[3]) // 2
        while cells[cy, cx] < 0:
            cx += 1
        return (cx, cy)

    def bfs(a, b):
        prev = {a: None}
        q = deque([a])
        while q:
            c = q.popleft()
            if c == b:
                break
            for d in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                nx, ny = c[0] + d[0], c[1] + d[1]
                if 0 <= nx < gw and 0 <= ny < gh and cells[ny, nx] >= 0 and (nx, ny) not in prev:
                    prev[(nx, ny)] = c
                    q.append((nx, ny))
        if b not in prev:
            return None
        out = []
        c = b
        while c is not None:
            out.append(c)
            c = prev[c]
        return out[::-1]

    stops = [centre(r) for r in gm.rooms] if len(gm.rooms) > 1 else None
    if stops is None:  # stress map: a lawnmower path over the open grid
        stops = []
        for k, cy in enumerate(range(2, gh - 2, 5)):
            xs = [cx for cx in range(2, gw - 2) if cells[cy, cx] >= 0]
            if xs:
                stops += [(xs[0], cy), (xs[-1], cy)] if k % 2 == 0 else [(xs[-1], cy), (xs[0], cy)]
    path = [stops[0]]
    for a, b in zip(stops[:-1], stops[1:]):
        seg = bfs(path[-1], b)
        if seg:
            path += seg[1:]
    pts = np.array([(gm.wx(cx) + gm.cell * 0.5 + 3.25, gm.wy(cy) + gm.cell * 0.5 - 5.5) for cx, cy in path], dtype=np.float64)
    # drop duplicate points, smooth nothing: headings are piecewise constant, the sin() term adds the sweep
    d = np.sqrt(((pts[1:] - pts[:-1]) ** 2).sum(1))
    keep = np.concatenate([[True], d > 0])
    pts = pts[keep]
    d = np.sqrt(((pts[1:] - pts[:-1]) ** 2).sum(1))
    s = np.concatenate([[0.0], np.cumsum(d)])
    t = (np.arange(n) + 0.37) * (s[-1] / n)
    idx = np.clip(np.searchsorted(s, t, side="right") - 1, 0, len(d) - 1)
    f = (t - s[idx]) / d[idx]
    xy = pts[idx] + (pts[idx + 1] - pts[idx]) * f[:, None]
    heading = np.arctan2(pts[idx + 1, 1] - pts[idx, 1], pts[idx + 1, 0] - pts[idx, 0])
    ang = heading + 0.6 * np.sin(2.0 * np.pi * np.arange(n) / 257.0)
    return np.stack([xy[:, 0], xy[:, 1], ang], axis=1).astype(np.float32)


def scatter_viewpoints(gm: GridMap, n: int, seed: int = SEED ^ 0x5EED) -> np.ndarray:
    """Config 5: uniform positions in cell interiors, uniform angle in [0, 2*pi)."""
    rng = PCG32(seed)
    ys, xs = np.nonzero(gm.cells >= 0)
    out = np.empty((n, 3), dtype=np.float32)
    for i in range(n):
        k = rng.below(len(xs))
        out[i, 0] = gm.wx(int(xs[k])) + 8 + (gm.cell - 16) * (rng.next() / 4294967296.0)
        out[i, 1] = gm.wy(int(ys[k])) + 8 + (gm.cell - 16) * (rng.next() / 4294967296.0)
        out[i, 2] = 2.0 * math.pi * (rng.next() / 4294967296.0)
    return out


if __name__ == "__main__":
    import sys
    kind = sys.argv[1] if len(sys.argv) > 1 else "e1m1"
    out = sys.argv[2] if len(sys.argv) > 2 else "synth_%s.wad" % kind
    wad, gm, stats = build_wad(kind)
    open(out, "wb").write(wad)
    print(out, len(wad), "bytes", stats)
