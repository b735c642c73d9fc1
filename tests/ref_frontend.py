"""Independent numpy-float32 restatement of the stateless part of the reference's FRONT-END, written from the Rust sources
(not from oracle/drr_oracle.cpp, not from csrc/): the pieces that decide WHICH arguments reach the leaf drawers.

    Vertex::rotate / is_left_of_line / distance_to      src/map/vertexes.rs:20-39
    Line::intersection                                   src/geometry.rs:54-80
    clip_to_viewport                                     src/renderer/misc.rs:13-115
    perspective_transform, make_sidedef_non_vertical_line  src/renderer/misc.rs:131-161
    process_seg: which parts a seg is cut into           src/renderer/segs.rs:348-590
    process_sidedef: screen x range, per-column bottom_y / top_y   src/renderer/segs.rs:128-208
    the map lumps themselves (own WAD reader)            src/map/{vertexes,linedefs,sidedefs,sectors,segs}.rs

Every operation is one IEEE f32 operation in the reference's order (numpy float32 scalars; cos/sin from the C library the
oracle links, because libm's last bit is not specified).  tests/test_oracle.py uses it to check the oracle's trace: every
wall call the oracle makes must carry, bit for bit, arguments this module derives for some part of some seg.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL("libm.so.6")
for _f in (_libm.cosf, _libm.sinf):
    _f.restype = ctypes.c_float
    _f.argtypes = [ctypes.c_float]

TWOSIDED, DONTPEGTOP, DONTPEGBOTTOM = 4, 8, 16  # src/map/linedefs.rs:12-14
PLAYER_EYE_HEIGHT = f32(41.0)                   # src/renderer/constants.rs:3


def rust_as_i32(x) -> int:
    """`f32 as i32`: truncate toward zero, saturate, NaN -> 0."""
    x = float(x)
    if x != x:
        return 0
    return int(max(-2147483648.0, min(2147483647.0, np.trunc(x))))


def rust_as_i16(x) -> int:
    x = float(x)
    if x != x:
        return 0
    return int(max(-32768.0, min(32767.0, np.trunc(x))))


def wrap_i16(v: int) -> int:
    return ((int(v) + 32768) & 0xFFFF) - 32768


class Constants:
    """src/renderer/constants.rs:7-17 for a W x H screen."""

    def __init__(self, W: int, H: int):
        self.W, self.H = W, H
        self.ASPECT = f32(200.0) / f32(240.0)
        self.GAME_SCREEN_WIDTH = f32(W) / self.ASPECT
        self.GCFX = self.GAME_SCREEN_WIDTH / f32(2.0)
        self.CFX = f32(W) / f32(2.0)
        self.CFY = f32(H) / f32(2.0)


# ---- vertexes.rs / geometry.rs ------------------------------------------------------------------------------------------
def rotate(v, angle):
    c, s = f32(_libm.cosf(float(angle))), f32(_libm.sinf(float(angle)))
    return (f32(v[0] * c - v[1] * s), f32(v[1] * c + v[0] * s))


def sub(a, b):
    return (f32(a[0] - b[0]), f32(a[1] - b[1]))


def cross(a, b):
    return f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))


def is_left_of_line(p, line):
    return bool(cross(sub(p, line[0]), sub(line[1], line[0])) <= f32(0.0))


def distance_to(a, b):
    dx, dy = f32(a[0] - b[0]), f32(a[1] - b[1])
    return f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy))))


def intersection(l1, l2):
    (x1, y1), (x2, y2) = l1
    (x3, y3), (x4, y4) = l2
    quot = f32(f32(f32(x1 - x2) * f32(y3 - y4)) - f32(f32(y1 - y2) * f32(x3 - x4)))
    if abs(quot) < f32(0.001):
        return None
    inv = f32(f32(1.0) / quot)
    a = f32(f32(x1 * y2) - f32(y1 * x2))
    b = f32(f32(x3 * y4) - f32(y3 * x4))
    px = f32(inv * f32(f32(a * f32(x3 - x4)) - f32(f32(x1 - x2) * b)))
    py = f32(inv * f32(f32(a * f32(y3 - y4)) - f32(f32(y1 - y2) * b)))
    return (px, py)


# ---- misc.rs ----------------------------------------------------------------------------------------------------------------
def clip_to_viewport(line):
    """misc.rs:13-115.  line = ((sx, sy), (ex, ey)) in player coordinates.  Returns (start, end, start_offset) or None."""
    z, o = f32(0.0), f32(1.0)
    left, right = ((z, z), (o, o)), ((z, z), (o, f32(-1.0)))
    start, end = line
    start_outside_left, end_outside_left = is_left_of_line(start, left), is_left_of_line(end, left)
    start_outside_right, end_outside_right = not is_left_of_line(start, right), not is_left_of_line(end, right)
    start_in = bool(start[0] > z) and not start_outside_left and not start_outside_right
    end_in = bool(end[0] > z) and not end_outside_left and not end_outside_right
    if start_in and end_in:
        return (start, end, f32(0.0))
    li, ri = intersection(line, left), intersection(line, right)
    left_hit = li is not None and bool(li[0] >= z)
    right_hit = ri is not None and bool(ri[0] >= z)
    if not start_in and not end_in and not left_hit and not right_hit:
        return None
    if not start_in and not end_in and (left_hit != right_hit):
        return None
    if (right_hit and start_outside_right and end_outside_right) or (left_hit and start_outside_left and end_outside_left):
        return None
    start_offset = f32(0.0)
    if left_hit:
        if start_outside_left:
            start_offset = distance_to(li, start)
            start = li
        if end_outside_left:
            end = li
    if right_hit:
        if start_outside_right:
            start = ri
        if end_outside_right:
            end = ri
    return (start, end, start_offset)


def make_sidedef_non_vertical_line(k: Constants, start, end, height):
    """misc.rs:131-161: ((x0, y0), (x1, y1)) integer screen points of the slanted edge at `height` (view space)."""
    def point(v):
        tx = f32(f32(k.GCFX * v[1]) / v[0])       # perspective_transform: x = v.y, z = v.x
        ty = f32(f32(k.GCFX * height) / v[0])
        tx = f32(tx * k.ASPECT)
        x = rust_as_i32(f32(k.CFX - tx))
        y = rust_as_i32(f32(k.CFY - ty))
        return (min(x, k.W - 1), y)
    return point(start), point(end)


# ---- the map lumps -----------------------------------------------------------------------------------------------------------
class MapLumps:
    """VERTEXES, LINEDEFS, SIDEDEFS, SECTORS, SEGS of one map, read straight from the WAD (src/wad.rs:86-109 directory)."""

    def __init__(self, path: str, map_name: str = "E1M1"):
        data = open(path, "rb").read()
        _, n, ofs = struct.unpack_from("<4sii", data, 0)
        entries = [struct.unpack_from("<ii8s", data, ofs + 16 * i) for i in range(n)]
        names = [e[2].rstrip(b"\0").decode("ascii") for e in entries]
        at = names.index(map_name)
        lump = {}
        for i in range(at + 1, min(at + 12, n)):
            lump.setdefault(names[i], data[entries[i][0]:entries[i][0] + entries[i][1]])

        def name8(b):
            return b.split(b"\0")[0].decode("ascii").upper()

        self.vertexes = [(f32(x), f32(y)) for x, y in struct.iter_unpack("<hh", lump["VERTEXES"])]
        self.sectors = [dict(floor=fl, ceil=ce, floor_tex=name8(ft), ceil_tex=name8(ct), light=li)
                        for fl, ce, ft, ct, li, _, _ in struct.iter_unpack("<hh8s8shhh", lump["SECTORS"])]
        self.sidedefs = [dict(xoff=xo, yoff=yo, upper=name8(u), lower=name8(lo), middle=name8(m), sector=s)
                         for xo, yo, u, lo, m, s in struct.iter_unpack("<hh8s8s8sh", lump["SIDEDEFS"])]
        self.linedefs = [dict(v1=a, v2=b, flags=fl, front=fr, back=bk) for a, b, fl, _, _, fr, bk in struct.iter_unpack("<hhhhhhh", lump["LINEDEFS"])]
        self.segs = [dict(v1=a, v2=b, linedef=ld, direction=d != 0, offset=off) for a, b, _, ld, d, off in struct.iter_unpack("<hhhhhh", lump["SEGS"])]


def seg_parts(k: Constants, m: MapLumps, seg, px, py, angle, floor_height):
    """process_seg (segs.rs:348-590) for one seg: the list of process_sidedef calls it makes, each as a dict of the arguments
    render_vertical_bitmap_line would get (clipped line, start_offset, start_x, end_x, bottom/top height, offsets, light level,
    texture name, flags) plus the per-column (bottom_y, top_y) function.  Empty when the seg is culled."""
    ld = m.linedefs[seg["linedef"]]
    front_i, back_i = (ld["back"], ld["front"]) if seg["direction"] else (ld["front"], ld["back"])
    if front_i < 0:
        return []
    front_sd = m.sidedefs[front_i]
    front = m.sectors[front_sd["sector"]]
    back = m.sectors[m.sidedefs[back_i]["sector"]] if back_i >= 0 else None
    floor_h, ceil_h = f32(front["floor"]), f32(front["ceil"])
    portal_bottom = f32(back["floor"]) if back is not None and back["floor"] > front["floor"] else None
    portal_top = f32(back["ceil"]) if back is not None and back["ceil"] < front["ceil"] else None
    two_sided = (ld["flags"] & TWOSIDED) != 0
    top_unpegged, bottom_unpegged = (ld["flags"] & DONTPEGTOP) != 0, (ld["flags"] & DONTPEGBOTTOM) != 0
    pos = (f32(px), f32(py))
    start = rotate(sub(m.vertexes[seg["v1"]], pos), f32(-f32(angle)))
    end = rotate(sub(m.vertexes[seg["v2"]], pos), f32(-f32(angle)))
    cl = clip_to_viewport((start, end))
    if cl is None:
        return []
    cs, ce, start_offset = cl
    player_height = f32(f32(floor_height) + PLAYER_EYE_HEIGHT)
    fl0, fl1 = make_sidedef_non_vertical_line(k, cs, ce, f32(floor_h - player_height))
    if fl0[0] > fl1[0]:
        return []  # facing the back of it
    draw_ceiling = True
    if back is not None and "SKY" in front["ceil_tex"] and "SKY" in back["ceil_tex"]:
        portal_top = None
        ceil_h = min(f32(back["ceil"]), ceil_h)
        draw_ceiling = False
    parts = []

    def part(bottom_height, top_height, offset_y, texture, flags):
        b0, b1 = make_sidedef_non_vertical_line(k, cs, ce, bottom_height)
        t0, t1 = make_sidedef_non_vertical_line(k, cs, ce, top_height)
        if wrap_i16(b0[0]) == wrap_i16(b1[0]) or wrap_i16(t0[0]) == wrap_i16(t1[0]):
            return  # looked at dead on from the side
        bottom_delta = f32(f32(f32(b0[1]) - f32(b1[1])) / f32(f32(b0[0]) - f32(b1[0])))
        top_delta = f32(f32(f32(t0[1]) - f32(t1[1])) / f32(f32(t0[0]) - f32(t1[0])))

        def column(x):
            by = rust_as_i16(f32(f32(b0[1]) + f32(f32(f32(x) - f32(b0[0])) * bottom_delta)))
            ty = rust_as_i16(f32(f32(t0[1]) + f32(f32(f32(x) - f32(t0[0])) * top_delta)))
            return by, ty
        parts.append(dict(line=(cs[0], cs[1], ce[0], ce[1]), start_offset=start_offset, start_x=b0[0], end_x=b1[0], bottom_height=bottom_height,
                          top_height=top_height, offset_x=wrap_i16(front_sd["xoff"] + seg["offset"]), offset_y=wrap_i16(front_sd["yoff"] + wrap_i16(offset_y)),
                          light_level=front["light"], texture=texture, flags=flags, column=column))

    if not two_sided:
        offset_y = rust_as_i32(f32(floor_h - ceil_h)) if bottom_unpegged else 0
        part(f32(floor_h - player_height), f32(ceil_h - player_height), offset_y, front_sd["middle"], "solid")
    else:
        part(f32(floor_h - player_height), f32(ceil_h - player_height), 0, front_sd["middle"], "occlusions")
        mid_floor = portal_bottom if portal_bottom is not None else floor_h
        mid_ceil = portal_top if portal_top is not None else ceil_h
        part(f32(mid_floor - player_height), f32(mid_ceil - player_height), 0, front_sd["middle"], "two_sided_middle")
        if portal_bottom is not None:
            offset_y = rust_as_i32(f32(ceil_h - portal_bottom)) if bottom_unpegged else 0
            part(f32(floor_h - player_height), f32(portal_bottom - player_height), offset_y, front_sd["lower"], "lower")
        if portal_top is not None:
            offset_y = 0 if top_unpegged else rust_as_i32(f32(portal_top - ceil_h))
            part(f32(portal_top - player_height), f32(ceil_h - player_height), offset_y, front_sd["upper"], "upper")
    return parts
