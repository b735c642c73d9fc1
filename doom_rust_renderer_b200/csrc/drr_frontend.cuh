// drr_frontend.cuh -- the reference's front-end for the DEVICE: BSP walk, seg transform /
// clip, occlusion arrays, visplane building and the map objects of one viewpoint, written so that it WRITES the draw lists
// of include/drr.h's device representation (ops / SegRec / ColRec / PlaneRec / (top, bottom) pairs) instead of drawing.
// One WARP runs one Renderer::render():
//   Renderer::render / render_node      src/renderer/mod.rs:69-104,118-136
//   Segs::process_seg / process_sidedef src/renderer/segs.rs:121-590
//   clip_to_viewport, projection        src/renderer/misc.rs:13-161
//   SidedefVisPlanes                    src/renderer/sidedef_visplanes.rs:41-84
//   sector lookup                       src/renderer/bsp.rs:9-44
//   draw_map_objects                    src/renderer/map_objects.rs:19-241
//   BitmapRender ordering predicates    src/renderer/bitmap_render.rs:101-188
// All four phases: A (walls), B (visplanes), C (map objects: sprite projection, per-sprite clip arrays, depth order,
// interleave with the masked mid-textures behind them) and D (the remaining masked mid-textures).
//
// The code is plain scalar C++ shared by nvcc (device: drr_frontend_kernel in drr_frontend.cu) and the host compiler (the
// CPU test harness drr_test_fe_emit_views_host, test infrastructure only), so that the list equality with the host
// front-end is checked in the CPU test suite and the device run only has to reproduce the same IEEE operations: every
// f32 expression keeps the reference's evaluation order; the translation unit is built with -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=false (no contraction, IEEE division / sqrt, denormals kept); cos/sin of the player angle are
// evaluated by the host libm and passed in (ViewIn), like drr_view.
//
// The same code runs in two modes: EMIT = true writes the lists (into per-view slabs in single-pass mode, at exact
// offsets in two-pass mode), EMIT = false only counts (the first pass of two-pass mode).  A viewpoint on which the
// reference would panic reports FE_PANIC and gets no frame.
#pragma once
#include "drr_device.cuh"

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#define FE_NOINLINE inline __host__ __device__ __noinline__
#else
#define FE_HD inline
#define FE_NOINLINE inline
#endif

namespace drr {
namespace fe {

// ---- the map, flattened (host side: drr_fe_upload_map) ----------------------------------------------------------
struct Node { // map/nodes.rs
    float x, y, dx, dy;
    int32_t right, left; // >= 0 node, < 0: ~subsector
};
struct SubSector {
    int32_t first, count;
};
struct Seg { // map/segs.rs; vertices resolved
    float v1x, v1y, v2x, v2y;
    int32_t line;
    int16_t dir, offset;
};
struct Line {
    int32_t front, back; // sidedef index, -1 none
    int32_t flags;
};
struct Side {
    float xoff, yoff;
    int32_t upper, lower, middle; // index into Map::bitmaps, -1 = "-", -2 = unknown texture name
    int32_t sector;
};
struct Sector { // flats resolved for the batch's timestamp (flats.rs:103-111)
    int16_t floor, ceil, light;
    int16_t ceil_name_has_sky;    // segs.rs:464-469 tests the sector's texture NAME
    int16_t floor_flat, ceil_flat; // flat SLOT of the context, -2 = lump missing (panics when shown)
    int16_t floor_sky, ceil_sky;   // the resolved flat's name contains "SKY" (visplanes.rs:89)
};
struct Bitmap {
    int32_t slot; // bitmap slot of the context, -1 = never uploaded (zero-sized)
    uint32_t base;
    int16_t w, h;
    uint32_t opaque;
};
struct Thing { // a non-null map object at tic 0 (map_objects.rs:25-50), everything view-independent resolved by the host
    float x, y, angle;
    int32_t sector;      // sector_at(position), -1 = outside the map (bsp.rs:9-44)
    int32_t full_bright; // info.rs state flag
    int32_t rotate;      // the sprite frame has 8 rotations
    int32_t bitmap[8];   // index into Map::bitmaps per rotation (entry 0 when !rotate)
    int16_t top_offset[8];
};
struct NodeUp { // the BSP tree seen from below (drr_fe_upload_map builds it when the node lump is a proper tree)
    int32_t up;          // (parent node << 1) | (1 when this is the parent's LEFT child), -1 at the root
    int32_t left, right; // segs under the left / right child (node entries only)
};
struct Map {
    const Node *nodes;
    const SubSector *ssectors;
    const NodeUp *node_up; // [nnodes], null when the lump is not a tree (the walk then runs serially with a stack)
    const int32_t *ss_up;  // [nssectors] NodeUp::up of every subsector
    int nssectors, nord;   // nord = segs the walk lists = sum of the subsectors' counts
    int side_words;        // (nnodes + 31) / 32: words of the per-view "player is left of the node's line" bit set
    const Seg *segs;
    const Line *lines;
    const Side *sides;
    const Sector *sectors;
    const Bitmap *bitmaps;
    const Thing *things;
    int nthings;
    int nnodes, nsegs;
    int W, H;
    float ASPECT, GCFX, CFX, CFY; // constants.rs:7-17, derived by drr_ctx_create
    int sky_kind;                 // KIND_SKY / KIND_SKY_HOLES, -1 = sky not set
    int phases;                   // DRR_PHASES_* (1 walls, 2 planes, 4 masked)
};

struct ViewIn { // player (game.rs:41-45) + host-evaluated cos/sin of angle and of -angle (Vertex::rotate, vertexes.rs:20-25)
    float x, y, angle;
    float cos_a, sin_a;   // cosf(angle), sinf(angle): shipped to the draw kernels in View
    float cos_n, sin_n;   // cosf(-angle), sinf(-angle): the view transform of segs.rs:380-383
};

enum : uint32_t { FE_OK = 0, FE_PANIC = 1, FE_HARD = 2 };
enum : uint32_t { // detail codes (messages: fe_detail_message() in drr_api.cu)
    FED_NONE = 0,
    FED_CLIP_X,          // "Clipped line x < -0.01"            segs.rs:389-391
    FED_UNKNOWN_TEXTURE, // "Unknown texture"                   segs.rs:152-157
    FED_NOT_VERTICAL,    // "Wall start not vertical"           segs.rs:159-167
    FED_LINE_X,          // "Invalid line start/end x"          segs.rs:169-184
    FED_FLAT_MISSING,    // flat lump missing                   flats.rs:92-100
    FED_STACK,           // BSP deeper than the walk stack, or subsectors that share segs (hard)
    FED_BITMAP_SLOT,     // emitted bitmap was never uploaded   (hard: DRR_E_ASSET on the host path)
    FED_SKY_UNSET,       // sky visplane but no sky bitmap      (hard: DRR_E_ASSET on the host path)
    FED_CAPACITY,        // a view's lists outgrew its slab     (single-pass mode only: the batch is redone with the count pass)
    FED_SCRATCH,         // a view outgrew the front-end's working arrays (hard)
    FED_ROTATION,        // "Invalid rotation"                  map_objects.rs:60-62
    FED_MO_X,            // map object column outside the screen map_objects.rs:170-176
};

struct Counts { // per viewpoint, written by the count pass. 40 bytes
    uint32_t nops, nsegs, ncols, nplanes, nparr, reccap;
    uint32_t status, detail;
    uint32_t nrec; // columns that survive clipping (what the bin kernel will actually write; statistics)
    uint32_t pad[3];
};
struct Caps { // how much room a view has (single-pass mode: its slab; two-pass mode: exactly what the count pass found)
    uint32_t ops, segs, cols, planes, parr;
};
struct Bases { // per viewpoint, written by the host between the passes. 32 bytes
    uint32_t op, seg, col, plane, parr;
    int32_t frame; // recorded frame index, -1 = no frame (panic)
    uint32_t pad[2];
};

// masked phase only: what the reference keeps per BitmapRender (bitmap_render.rs:29-45) and per visible map object
struct RenderRec { // 36 bytes
    float lsx, lsy, lex, ley; // BitmapRender.clipped_line.line
    uint32_t col0, ncol;      // its columns in Scratch::allcols
    int32_t dseg;             // its header in Scratch::dsegs (-1: nothing to draw: texture "-")
    uint32_t flags;           // RF_*
    int16_t x0, x1;           // x of its first / last column
};
enum : uint32_t { RF_TWOSIDED = 1, RF_EXT_BOTTOM = 2, RF_EXT_TOP = 4, RF_DRAW_CEILING = 8, RF_DRAWN = 16 };
struct MoRec { // 16 bytes
    float vx, vy;  // midpoint of the clipped sprite line (map_objects.rs:222-225)
    int32_t key;   // `line.start.x as i16` (map_objects.rs:216)
    int32_t dseg;  // its header in Scratch::dsegs (-1: no column)
};

struct SegPre;
// per-viewpoint scratch (global memory)
struct Scratch {
    uint8_t *hor_ocl;              // W entries each: segs.rs:97-99
    int16_t *floor_ocl, *ceil_ocl;
    uint32_t *rows[2]; // W (top, bottom) pairs of the visplane being accumulated: 0 = bottom (floor), 1 = top (ceiling)
    int32_t *order;    // nsegs entries: the map's segs in this view's BSP order
    uint32_t *side;    // Map::side_words words: bit i = the player is on the left of node i's partition line
    const SegPre *pre;       // nsegs records of drr_fe_pre_kernel for this view (valid where the code is not 0), or null
    const uint8_t *pre_code; // nsegs codes
    // masked phase only
    RenderRec *renders; // the parts that can clip sprites or are drawn late, in creation order
    ColRec *allcols;    // their columns, and the sprites' columns
    SegRec *dsegs;      // headers of what is drawn late (masked mid-textures, sprites), in creation order
    MoRec *mos;         // the visible map objects
    int32_t *mo_order;  // their draw order
    int32_t *dseg_part; // cap_dsegs: the part (index into renders) of every masked mid-texture header in dsegs
    uint32_t cap_renders, cap_allcols, cap_dsegs, cap_mos;
};

struct Out { // where the emit pass writes (pointers are the batch's arrays; indices are global)
    View *views;
    uint32_t *ops;
    SegRec *segs;
    ColRec *cols;
    PlaneRec *planes;
    uint32_t *parr;
};

// ---- Rust scalar semantics (the same helpers as csrc/host/drr_scene.cpp) ------------------------------------------
// (on the device one conversion instruction each: cvt.rzi saturates and turns NaN into 0, which is Rust's `as`)
FE_HD int16_t as_i16(float f) {
#if defined(__CUDA_ARCH__)
    return (int16_t)sat_i16(f);
#endif
    if (f != f) return 0;
    if (f <= -32768.0f) return (int16_t)-32768;
    if (f >= 32767.0f) return (int16_t)32767;
    return (int16_t)f;
}
FE_HD int32_t as_i32(float f) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);
#endif
    if (f != f) return 0;
    if (f <= -2147483648.0f) return (int32_t)0x80000000;
    if (f >= 2147483648.0f) return 0x7fffffff;
    return (int32_t)f;
}
FE_HD uint8_t as_u8(float f) {
#if defined(__CUDA_ARCH__)
    return (uint8_t)sat_u8(f);
#endif
    if (f != f || f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
FE_HD int16_t w16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }

struct V2 {
    float x, y;
};
struct Seg2 {
    V2 s, e;
};
FE_HD V2 sub(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
FE_HD V2 rot(V2 v, float c, float s) { return V2{v.x * c - v.y * s, v.y * c + v.x * s}; } // vertexes.rs:20-25 with host cos/sin
FE_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
FE_HD bool left_of(V2 v, V2 ls, V2 le) { return cross(sub(v, ls), sub(le, ls)) <= 0.0f; } // vertexes.rs:32-34
FE_HD float dist(V2 a, V2 b) {
    const float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
FE_HD bool intersect(const Seg2 &a, const Seg2 &b, V2 *out) { // geometry.rs:56-82
    const float x1 = a.s.x, y1 = a.s.y, x2 = a.e.x, y2 = a.e.y, x3 = b.s.x, y3 = b.s.y, x4 = b.e.x, y4 = b.e.y;
    const float quot = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (fabsf(quot) < 0.001f) return false;
    const float inv = 1.0f / quot;
    out->x = inv * ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4));
    out->y = inv * ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4));
    return true;
}

struct ScreenLine {
    int32_t sx, sy, ex, ey;
};
// misc.rs:130-161 (perspective projection of a view-space line at one height).  The screen x of the two ends depends on
// the line only, the screen y on the line and the height: a seg's parts share the x pair.
struct ScreenX {
    int32_t sx, ex;
};
FE_HD ScreenX project_x(const Map &m, const Seg2 &l) {
    float tsx = m.GCFX * l.s.y / l.s.x, tex = m.GCFX * l.e.y / l.e.x;
    tsx *= m.ASPECT;
    tex *= m.ASPECT;
    ScreenX r{as_i32(m.CFX - tsx), as_i32(m.CFX - tex)};
    r.sx = r.sx < m.W - 1 ? r.sx : m.W - 1;
    r.ex = r.ex < m.W - 1 ? r.ex : m.W - 1;
    return r;
}
FE_HD ScreenLine project(const Map &m, const Seg2 &l, ScreenX sx, float height) {
    const float tsy = m.GCFX * height / l.s.x, tey = m.GCFX * height / l.e.x;
    return ScreenLine{sx.sx, as_i32(m.CFY - tsy), sx.ex, as_i32(m.CFY - tey)};
}

FE_NOINLINE bool clip_fov(const Seg2 &line, Seg2 *out, float *start_offset) { // misc.rs:13-115
    const V2 O = {0.0f, 0.0f}, LE = {1.0f, 1.0f}, RE = {1.0f, -1.0f};
    const Seg2 L = {O, LE}, R = {O, RE};
    const bool s_out_l = left_of(line.s, O, LE), e_out_l = left_of(line.e, O, LE);
    const bool s_out_r = !left_of(line.s, O, RE), e_out_r = !left_of(line.e, O, RE);
    const bool s_in = line.s.x > 0.0f && !s_out_l && !s_out_r;
    const bool e_in = line.e.x > 0.0f && !e_out_l && !e_out_r;
    if (s_in && e_in) {
        *out = line;
        *start_offset = 0.0f;
        return true;
    }
    V2 li = {0.0f, 0.0f}, ri = {0.0f, 0.0f};
    const bool lhit = intersect(line, L, &li) && li.x >= 0.0f;
    const bool rhit = intersect(line, R, &ri) && ri.x >= 0.0f;
    if (!s_in && !e_in && !lhit && !rhit) return false;
    if (!s_in && !e_in && lhit != rhit) return false;
    if ((rhit && s_out_r && e_out_r) || (lhit && s_out_l && e_out_l)) return false;
    V2 s = line.s, e = line.e;
    float so = 0.0f;
    if (lhit) {
        if (s_out_l) {
            so = dist(li, s);
            s = li;
        }
        if (e_out_l) e = li;
    }
    if (rhit) {
        if (s_out_r) s = ri;
        if (e_out_r) e = ri;
    }
    out->s = s;
    out->e = e;
    *start_offset = so;
    return true;
}

FE_NOINLINE int sector_at(const Map &m, V2 p) { // renderer/bsp.rs:9-44
    int n = m.nnodes - 1;
    for (;;) {
        const Node nd = m.nodes[n];
        const V2 a = {nd.x, nd.y}, b = {nd.x + nd.dx, nd.y + nd.dy};
        const int ch = left_of(p, a, b) ? nd.left : nd.right;
        if (ch >= 0) {
            n = ch;
            continue;
        }
        const SubSector ss = m.ssectors[~ch];
        for (int i = 0; i < ss.count; i++) {
            const Seg sg = m.segs[ss.first + i];
            const Line ld = m.lines[sg.line];
            const int sd = sg.dir ? ld.back : ld.front;
            if (sd != -1) return m.sides[sd].sector;
        }
        return -1;
    }
}

// The part of process_seg that touches no per-view state (segs.rs:353-460): the view transform, the clip against the
// field of view, the screen x of the ends, the back-face test.  On the device it runs for every (viewpoint, seg) pair in a
// kernel of its own (drr_fe_pre_kernel: one thread per pair, no order, no state), which leaves one 32-byte record per pair
// that survives and a code byte per pair; the per-view walk then only looks the codes up in BSP order.  Without those arrays
// (CPU test harness) the walk evaluates it in place, 32 segs at a time.
struct SegPre { // 32 bytes
    float csx, csy, cex, cey, so; // ClippedLine
    int32_t sx, ex;               // screen x of its ends
    int32_t code;                 // 0 draws nothing, 1 go on, 2 panics ("Clipped line x < -0.01"), 4 go on and nothing in its
                                  // sidedef can panic (no unknown texture, no missing flat): may be skipped when occluded
};
FE_NOINLINE SegPre seg_pre_of(const Map &m, V2 ppos, float cos_n, float sin_n, const Seg &sg) {
    SegPre p{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0, 0, 0};
    const Line ld = m.lines[sg.line];
    const int fi = sg.dir ? ld.back : ld.front;
    if (fi == -1) return p;
    const V2 v1 = {sg.v1x, sg.v1y}, v2 = {sg.v2x, sg.v2y};
    const Seg2 view = {rot(sub(v1, ppos), cos_n, sin_n), rot(sub(v2, ppos), cos_n, sin_n)};
    Seg2 cl;
    float so;
    if (!clip_fov(view, &cl, &so)) return p;
    p.csx = cl.s.x;
    p.csy = cl.s.y;
    p.cex = cl.e.x;
    p.cey = cl.e.y;
    p.so = so;
    if (cl.s.x < -0.01f) {
        p.code = 2;
        return p;
    }
    const ScreenX sx = project_x(m, cl);
    p.sx = sx.sx;
    p.ex = sx.ex;
    p.code = sx.sx > sx.ex ? 0 : 1; // back faces draw nothing
    if (p.code == 1 && sx.sx >= 0) {
        const Side sd = m.sides[fi];
        const Sector sec = m.sectors[sd.sector];
        if (sd.upper != -2 && sd.lower != -2 && sd.middle != -2 && sec.floor_flat >= 0 && sec.ceil_flat >= 0) p.code = 4;
    }
    return p;
}

// ---- one viewpoint = one warp ---------------------------------------------------------------------------------------
// The walk of a viewpoint is sequential from seg to seg (occlusion arrays, open visplanes), but inside a seg the screen
// columns are independent and inside a subsector so are the segs' visibility tests.  A WARP runs a viewpoint: everything
// that is per-view or per-seg is computed redundantly by all 32 lanes (uniform, no divergence), the column loop of
// process_sidedef gives lane l the column x0 + l of each 32-column chunk, and the order-dependent parts (column records
// are appended in x order; a visplane runs from its first point to the next flush) are resolved from ballots.
// On the host (CPU test harness) the 32 lanes are emulated by loops: FE_LANES(l) { ... } runs its body once per lane.
#if defined(__CUDA_ARCH__)
// (the __syncwarp at the end of a per-lane block: the per-view state lives in ONE copy per warp in shared memory, so the uniform
// code after the block must find every lane back in step before it reads and updates that state)
#define FE_LANES(l) for (int l = (int)(threadIdx.x & 31u), l##_1 = 1; l##_1; l##_1 = 0, __syncwarp())
#define FE_SYNC() __syncwarp()
#define FE_LEADER for (int ld_1 = 1; ld_1; ld_1 = 0, __syncwarp()) if ((threadIdx.x & 31u) == 0u)
template <class T>
struct PerLane { // a value every lane holds its own copy of
    T v;
    __device__ __forceinline__ T &operator[](int) { return v; }
};
#else
#define FE_LANES(l) for (int l = 0; l < 32; l++)
#define FE_SYNC() ((void)0)
#define FE_LEADER
template <class T>
struct PerLane {
    T v[32];
    T &operator[](int l) { return v[l]; }
};
#endif
template <class T>
struct LaneArr { // one value per lane inside the per-view state (which is one object per warp)
    T v[32];
    FE_HD T &operator[](int l) { return v[l]; }
};
template <class F>
FE_HD uint32_t ballot(F f) { // bit l = f(l)
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xffffffffu, f((int)(threadIdx.x & 31u)));
#else
    uint32_t mask = 0;
    for (int l = 0; l < 32; l++)
        if (f(l)) mask |= 1u << l;
    return mask;
#endif
}
template <class T>
FE_HD T from_lane(PerLane<T> &v, int src) { // lane src's value, on every lane (T: 32-bit scalar)
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xffffffffu, v.v, src);
#else
    return v[src];
#endif
}
FE_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#endif
    int c = 0;
    for (; v; v &= v - 1) c++;
    return c;
}
FE_HD int lowest(uint32_t v) { // index of the lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#endif
    int i = 0;
    while (!((v >> i) & 1u)) i++;
    return i;
}
FE_HD int highest(uint32_t v) { // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#endif
    int i = 31;
    while (!((v >> i) & 1u)) i--;
    return i;
}
FE_HD uint32_t below(int l) { // bits of the lanes < l, 0 <= l <= 32
#if defined(__CUDA_ARCH__)
    return l >= 32 ? 0xffffffffu : (1u << l) - 1u;
#endif
    return (uint32_t)((1ull << l) - 1ull);
}
FE_HD void prefetch_l1(const void *p) { // a hint: the line is on its way while the warp works on something else
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

template <bool EMIT>
struct Frame {
    const Map &m;
    Scratch sc;
    Out out;
    Bases base;   // EMIT only
    Counts n;     // running counts == cursors relative to the bases (uniform over the warp)
    Caps cap;     // EMIT only
    V2 ppos;
    float cos_n, sin_n, pfloor;
    // SidedefVisPlanes of the sidedef part being processed
    bool open[2];
    int16_t pl_left[2], pl_right[2];
    int16_t pl_flat[2], pl_sky[2], pl_height[2], pl_light;
    // Masked phase: parts that can clip sprites or are drawn late are remembered (Scratch::renders / allcols / dsegs) and
    // drawn -- appended to the output lists -- when the reference draws them (phases C and D at the end of run()).
    uint32_t nrenders, nallcols, ndsegs, nmos;
    uint32_t nmids; // masked mid-texture headers: dsegs[0 .. nmids), known when the walk is over
    // Which screen columns are fully occluded (hor_ocl), as a bit mask spread over the lanes: word w (columns 32w .. 32w+31)
    // lives on lane w % 32, slot w / 32.  A seg all of whose columns are occluded can only re-occlude them and flush
    // visplanes that are not open (segs.rs:186-330): it is skipped without touching memory.
    LaneArr<uint32_t> occ[4];
    bool occ_on; // W <= 4096

    FE_HD Frame(const Map &map) : m(map) {}

    FE_HD void fail(uint32_t status, uint32_t detail) {
        if (n.status == FE_OK) {
            n.status = status;
            n.detail = detail;
        }
    }

    // flush, sidedef_visplanes.rs:41-58: pushes the bottom visplane, then the top one.  mod.rs:106-116 later draws the
    // visplanes in push order (phase B comes after every wall), so the op words are appended after the walk; here the
    // plane's record is written and its rows [left, right] are copied from the accumulation buffer (the reference keeps
    // zero-initialised [i16; W] arrays, visplanes.rs:36-37: columns inside the range that never got a point are (0, 0)).
    FE_NOINLINE void flush() {
        if (!open[0] && !open[1]) return;
        for (int which = 0; which < 2; which++) {
            if (!open[which]) continue;
            open[which] = false;
            if (!(m.phases & 2)) continue;
            const int left = pl_left[which], right = pl_right[which];
            const uint32_t ncols = (uint32_t)(right - left + 1);
            const bool sky = pl_sky[which] != 0;
            if (sky && m.sky_kind < 0) fail(FE_HARD, FED_SKY_UNSET);
            const uint32_t arr_first = base.parr + n.nparr;
            const uint32_t *src = sc.rows[which] + left;
            if (EMIT && (n.nplanes + 1 > cap.planes || n.nparr + ncols > cap.parr)) {
                fail(FE_HARD, FED_CAPACITY);
                continue;
            }
            const int H = m.H;
            for (uint32_t i0 = 0; i0 < ncols; i0 += 32) {
                // how many of the columns draw anything (drr_api.cu: rec_emit_visplane)
                n.nrec += (uint32_t)popc32(ballot([&](int l) {
                    if (i0 + (uint32_t)l >= ncols) return false;
                    const uint32_t tb = src[i0 + l];
                    const int t = (int)(int16_t)(tb & 0xffffu) > 0 ? (int)(int16_t)(tb & 0xffffu) : 0; // visplanes.rs:61 / :95
                    const int b = (int)(int16_t)(tb >> 16) < H - 1 ? (int)(int16_t)(tb >> 16) : H - 1;   // :62 / :96
                    if (!sky && (int16_t)(b - t) <= 1) return false;                                     // :99-101 (not applied to sky)
                    return t <= b;
                }));
                if (EMIT) {
                    FE_LANES(l) {
                        if (i0 + (uint32_t)l < ncols) out.parr[arr_first + i0 + l] = src[i0 + l];
                    }
                }
            }
            if (EMIT) {
                FE_LEADER {
                    PlaneRec p;
                    p.flat_slot = sky ? (int16_t)-1 : pl_flat[which];
                    p.height = pl_height[which];
                    p.light_level = pl_light;
                    p.left = (int16_t)left;
                    p.right = (int16_t)right;
                    p.kind = (int16_t)(sky ? m.sky_kind : (int)KIND_FLAT);
                    p.arr_first = arr_first;
                    out.planes[base.plane + n.nplanes] = p;
                }
            }
            n.nplanes++;
            n.nparr += ncols;
            n.reccap += ncols;
        }
        FE_SYNC(); // the accumulation buffers may be written again
    }

    FE_HD void occlude(int x) { // segs.rs:113-117
        sc.hor_ocl[x] = 1;
        sc.floor_ocl[x] = (int16_t)(m.H / 2);
        sc.ceil_ocl[x] = (int16_t)(m.H / 2);
    }

    // process_sidedef, segs.rs:121-350
    FE_NOINLINE void sidedef_part(const Seg2 &cl, ScreenX sx, float start_offset, const Side &sd, int16_t seg_offset, const Sector &sec, float bottom_h,
                            float top_h, int32_t offset_y, int tex, bool only_occ, bool lower, bool upper, bool draw_ceiling, bool two_sided_mid) {
        const ScreenLine bottom = project(m, cl, sx, bottom_h), top = project(m, cl, sx, top_h);
        if (tex == -2) return fail(FE_PANIC, FED_UNKNOWN_TEXTURE);
        if (bottom.sx != top.sx || bottom.ex != top.ex) return fail(FE_PANIC, FED_NOT_VERTICAL);
        if ((int16_t)bottom.sx == (int16_t)bottom.ex || (int16_t)top.sx == (int16_t)top.ex) return;
        if (bottom.sx < 0 || bottom.sx >= m.W || bottom.ex < 0 || bottom.ex >= m.W) return fail(FE_PANIC, FED_LINE_X);
        const float bottom_delta = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
        const float top_delta = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);

        open[0] = open[1] = false;
        pl_flat[0] = sec.floor_flat;
        pl_flat[1] = sec.ceil_flat;
        pl_sky[0] = sec.floor_sky;
        pl_sky[1] = sec.ceil_sky;
        pl_height[0] = sec.floor;
        pl_height[1] = sec.ceil;
        pl_light = sec.light;
        const bool full_height = !lower && !upper && !only_occ;
        const int16_t H16 = (int16_t)m.H, Hm1 = w16(H16 - 1);
        // is this part's column list ever emitted?  phase A walls are drawn at once (segs.rs:231-258), two-sided middle
        // textures are deferred (phase D, segs.rs:593-597); occlusion-only parts and "-" textures draw nothing
        const bool wall = !two_sided_mid && !only_occ && (m.phases & 1) && tex >= 0;       // drawn now: its columns go to the output
        const bool track = (m.phases & 4) && !only_occ;                                     // remembered: clips sprites / drawn late
        const bool deferred = two_sided_mid && (m.phases & 4) && tex >= 0;
        const bool keep = wall || track;
        const bool planes_here = !two_sided_mid && (full_height || only_occ);
        const uint32_t col0 = n.ncols, acol0 = nallcols;
        uint32_t ncol = 0;
        int x_first = 0, x_last = 0;
        const int xe = bottom.ex; // the reference's `for x in start.x..end.x + 1`

        for (int c0 = bottom.sx; c0 <= xe; c0 += 32) {
            // ---- per column (lane l: x = c0 + l): segs.rs:186-330
            enum : uint32_t { EV_P0 = 1, EV_P1 = 2, EV_FLUSH = 4, EV_COL = 8, EV_OCC = 16 };
            PerLane<uint32_t> ev, row0, row1;
            PerLane<ColRec> col;
            FE_LANES(l) {
                const int x = c0 + l;
                uint32_t e = 0;
                if (x <= xe) {
                    if (!sc.hor_ocl[x]) {
                        const int16_t bottom_y = as_i16((float)bottom.sy + ((float)(int16_t)x - (float)bottom.sx) * bottom_delta);
                        const int16_t top_y = as_i16((float)top.sy + ((float)(int16_t)x - (float)top.sx) * top_delta);
                        const int16_t fvo = sc.floor_ocl[x], cvo = sc.ceil_ocl[x];
                        int16_t cb = fvo < bottom_y ? fvo : bottom_y, ct = cvo > top_y ? cvo : top_y;
                        cb = Hm1 < cb ? Hm1 : cb;
                        ct = ct < 0 ? (int16_t)0 : ct;
                        const bool in_area = cb >= ct;
                        if (in_area && keep) { // add_column, bitmap_render.rs:84-99
                            e |= EV_COL;
                            ColRec c;
                            c.x = (int16_t)x;
                            c.clipped_top_y = ct;
                            c.clipped_bottom_y = cb;
                            c.bottom_y = bottom_y;
                            c.top_y = top_y;
                            col[l] = c;
                        }
                        if (planes_here && in_area) {
                            if (cb < fvo && cb != Hm1) { // add_bottom_point(x, cb, fvo)
                                e |= EV_P0;
                                row0[l] = (uint32_t)(uint16_t)cb | ((uint32_t)(uint16_t)fvo << 16);
                            }
                            if (draw_ceiling && ct > cvo && ct != -1) { // add_top_point(x, cvo, ct)
                                e |= EV_P1;
                                row1[l] = (uint32_t)(uint16_t)cvo | ((uint32_t)(uint16_t)ct << 16);
                            }
                            if (!(e & (EV_P0 | EV_P1))) e |= EV_FLUSH;
                        } else if (planes_here && !in_area && fvo > cvo) {
                            if (bottom_y <= cvo) {
                                e |= EV_P0 | EV_OCC;
                                row0[l] = (uint32_t)(uint16_t)cvo | ((uint32_t)(uint16_t)fvo << 16);
                                occlude(x);
                            }
                            if (draw_ceiling && top_y >= fvo) {
                                e |= EV_P1 | EV_OCC;
                                row1[l] = (uint32_t)(uint16_t)cvo | ((uint32_t)(uint16_t)fvo << 16);
                                occlude(x);
                            }
                        }
                        if (!two_sided_mid && in_area && only_occ) {
                            sc.floor_ocl[x] = cb;
                            if (draw_ceiling) sc.ceil_ocl[x] = ct;
                        }
                        if (!two_sided_mid && in_area && lower) sc.floor_ocl[x] = ct;
                        if (!two_sided_mid && in_area && upper) sc.ceil_ocl[x] = cb;
                    } else if (planes_here) {
                        e |= EV_FLUSH;
                    }
                    if (!two_sided_mid && full_height) {
                        occlude(x);
                        e |= EV_OCC;
                    }
                }
                ev[l] = e;
            }
            if (occ_on && !two_sided_mid) { // newly occluded columns -> the lanes that own their mask words
                const uint32_t m_occ = ballot([&](int l) { return (ev[l] & EV_OCC) != 0; });
                if (m_occ) {
                    const int w0 = c0 >> 5, sh = c0 & 31;
                    const uint32_t lo_bits = m_occ << sh, hi_bits = sh ? m_occ >> (32 - sh) : 0u;
                    FE_LANES(l) { // (constant slot indices: the masks stay in registers)
                        const uint32_t a = (w0 & 31) == l ? lo_bits : 0u, b = ((w0 + 1) & 31) == l ? hi_bits : 0u;
                        const int sa = (w0 >> 5) & 3, sb = ((w0 + 1) >> 5) & 3;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                        for (int s4 = 0; s4 < 4; s4++) occ[s4][l] |= (sa == s4 ? a : 0u) | (sb == s4 ? b : 0u);
                    }
                }
            }
            uint32_t m_p0 = 0, m_p1 = 0, m_fl = 0, m_col = 0;
            if (planes_here) {
                m_p0 = ballot([&](int l) { return (ev[l] & EV_P0) != 0; });
                m_p1 = ballot([&](int l) { return (ev[l] & EV_P1) != 0; });
                m_fl = ballot([&](int l) { return (ev[l] & EV_FLUSH) != 0; });
            }
            if (keep) m_col = ballot([&](int l) { return (ev[l] & EV_COL) != 0; });
            // ---- column records, in x order
            if (m_col) {
                const uint32_t cnt = (uint32_t)popc32(m_col);
                if (wall && EMIT && n.ncols + cnt > cap.cols) {
                    fail(FE_HARD, FED_CAPACITY);
                    m_col = 0;
                } else if (track && nallcols + cnt > sc.cap_allcols) {
                    fail(FE_HARD, FED_SCRATCH);
                    m_col = 0;
                }
            }
            if (m_col) {
                const uint32_t cnt = (uint32_t)popc32(m_col);
                FE_LANES(l) {
                    if (ev[l] & EV_COL) {
                        const uint32_t k = (uint32_t)popc32(m_col & below(l));
                        if (wall && EMIT) out.cols[base.col + n.ncols + k] = col[l];
                        if (track) sc.allcols[nallcols + k] = col[l];
                    }
                }
                if (ncol == 0) x_first = c0 + lowest(m_col);
                x_last = c0 + highest(m_col);
                ncol += cnt;
                if (wall) {
                    n.ncols += cnt;
                    n.nrec += cnt; // 0 <= x < W and max(ct, 0) <= min(cb, H - 1) hold for every record
                }
                if (track) nallcols += cnt;
            }
            // ---- visplane rows: a lane inside an open visplane writes its point, or (0, 0) when it has none.  Visplane
            // `which` is open at lane l when it was open at the chunk's start and no lane below l flushes, or when a lane
            // between the last flush below l and l itself has a point.
            if (m_p0 | m_p1 | (uint32_t)(open[0] || open[1])) {
                FE_LANES(l) {
                    const int x = c0 + l;
                    if (x <= xe) {
                        const uint32_t fb = m_fl & below(l);
                        const int start = fb ? highest(fb) + 1 : 0;
                        const uint32_t run = below(l + 1) & ~below(start); // lanes start .. l
                        if (ev[l] & EV_P0) sc.rows[0][x] = row0[l];
                        else if ((!fb && open[0]) || (m_p0 & run)) sc.rows[0][x] = 0u;
                        if (ev[l] & EV_P1) sc.rows[1][x] = row1[l];
                        else if ((!fb && open[1]) || (m_p1 & run)) sc.rows[1][x] = 0u;
                    }
                }
                FE_SYNC(); // a flush below copies rows other lanes wrote
            }
            // ---- runs: points between two flushes belong to one pair of visplanes (sidedef_visplanes.rs:41-84)
            uint32_t f = m_fl;
            int lo = 0;
            if (!(m_p0 | m_p1) && !open[0] && !open[1]) continue; // nothing open, nothing opens: the flushes are no-ops
            for (;;) {
                const int nf = f ? lowest(f) : 32;
                const uint32_t run = below(nf) & ~below(lo); // lanes lo .. nf - 1
                const uint32_t pm[2] = {m_p0 & run, m_p1 & run};
                for (int which = 0; which < 2; which++)
                    if (pm[which]) {
                        if (!open[which]) {
                            open[which] = true;
                            pl_left[which] = (int16_t)(c0 + lowest(pm[which]));
                        }
                        pl_right[which] = (int16_t)(c0 + highest(pm[which]));
                    }
                if (nf == 32) break;
                if (open[0] || open[1]) flush();
                f &= f - 1;
                lo = nf + 1;
            }
        }
        FE_SYNC(); // the occlusion arrays written above are read by other lanes in the next part (its chunks start at another x)
        if (open[0] || open[1]) flush();
        if (!keep || ncol == 0 || n.status != FE_OK) return;
        SegRec r;
        if (wall || deferred) {
            const Bitmap bm = m.bitmaps[tex];
            if (bm.slot < 0) return fail(FE_HARD, FED_BITMAP_SLOT);
            r.bitmap_slot = (uint32_t)bm.slot;
            r.light_level = sec.light;
            r.phase = (int16_t)(deferred ? 2 : 0); // DRR_PHASE_MASKED / DRR_PHASE_WALL
            r.lsx = cl.s.x;
            r.lsy = cl.s.y;
            r.lex = cl.e.x;
            r.ley = cl.e.y;
            r.start_offset = start_offset;
            r.start_x = bottom.sx;
            r.end_x = bottom.ex;
            r.bottom_height = bottom_h;
            r.top_height = top_h;
            r.offset_x = w16(as_i16(sd.xoff) + seg_offset);
            r.offset_y = w16(as_i16(sd.yoff) + w16(offset_y));
            r.cols_first = base.col + col0;
            r.n = ncol;
            r.x0 = (int16_t)x_first;
            r.x1 = (int16_t)x_last;
            r.tex_base = bm.base;
            r.tex_w = bm.w;
            r.tex_h = bm.h;
            r.tex_opaque = bm.opaque;
            r.pad[0] = r.pad[1] = 0;
        }
        if (wall) { // segs.rs:231-258: drawn at once
            if (EMIT && (n.nsegs + 1 > cap.segs || n.nops + 1 > cap.ops)) return fail(FE_HARD, FED_CAPACITY);
            if (EMIT) {
                FE_LEADER {
                    out.segs[base.seg + n.nsegs] = r;
                    out.ops[base.op + n.nops] = base.seg + n.nsegs;
                }
            }
            n.nops++;
            n.nsegs++;
            // record slots the bin kernel may reserve: one per screen column inside the x range (drr_api.cu: rec_emit_columns)
            const int lo = x_first > 0 ? x_first : 0, hi = x_last < m.W - 1 ? x_last : m.W - 1;
            n.reccap += (uint32_t)(hi - lo + 1 > 0 ? hi - lo + 1 : 0);
        }
        if (track) { // Segs::bitmap_renders.push, segs.rs:332-349
            if (nrenders + 1 > sc.cap_renders || (deferred && ndsegs + 1 > sc.cap_dsegs)) return fail(FE_HARD, FED_SCRATCH);
            FE_LEADER {
                RenderRec rr;
                rr.lsx = cl.s.x;
                rr.lsy = cl.s.y;
                rr.lex = cl.e.x;
                rr.ley = cl.e.y;
                rr.col0 = acol0;
                rr.ncol = ncol;
                rr.dseg = deferred ? (int32_t)ndsegs : -1;
                rr.x0 = (int16_t)x_first;
                rr.x1 = (int16_t)x_last;
                rr.flags = (two_sided_mid ? RF_TWOSIDED : 0u) | ((lower || (!two_sided_mid && full_height)) ? RF_EXT_BOTTOM : 0u) |
                           ((upper || (!two_sided_mid && full_height)) ? RF_EXT_TOP : 0u) | (draw_ceiling ? RF_DRAW_CEILING : 0u);
                sc.renders[nrenders] = rr;
                if (deferred) {
                    r.cols_first = acol0; // for now: where its columns wait in allcols
                    sc.dsegs[ndsegs] = r;
                    sc.dseg_part[ndsegs] = (int32_t)nrenders;
                }
            }
            nrenders++;
            if (deferred) ndsegs++;
        }
    }

    // BitmapRender::render for what is drawn late (bitmap_render.rs:101-135): header dsegs[d] and its columns move to the
    // end of the view's output lists, exactly where the host front-end's drr_emit_columns call puts them.
    FE_NOINLINE void draw_late(int32_t d) {
        SegRec r = sc.dsegs[d];
        const uint32_t ncol = r.n, src = r.cols_first;
        if (EMIT && (n.nsegs + 1 > cap.segs || n.nops + 1 > cap.ops || n.ncols + ncol > cap.cols)) return fail(FE_HARD, FED_CAPACITY);
        const int W = m.W, H = m.H;
        for (uint32_t i0 = 0; i0 < ncol; i0 += 32) {
            PerLane<ColRec> c;
            FE_LANES(l) {
                if (i0 + (uint32_t)l < ncol) {
                    c[l] = sc.allcols[src + i0 + l];
                    if (EMIT) out.cols[base.col + n.ncols + i0 + l] = c[l];
                }
            }
            n.nrec += (uint32_t)popc32(ballot([&](int l) { // drr_api.cu: rec_emit_columns
                if (i0 + (uint32_t)l >= ncol) return false;
                const ColRec &k = c[l];
                if (k.x < 0 || k.x >= W) return false;
                return (k.clipped_top_y > 0 ? (int)k.clipped_top_y : 0) <= ((int)k.clipped_bottom_y < H - 1 ? (int)k.clipped_bottom_y : H - 1);
            }));
        }
        if (EMIT) {
            FE_LEADER {
                r.cols_first = base.col + n.ncols;
                out.segs[base.seg + n.nsegs] = r;
                out.ops[base.op + n.nops] = base.seg + n.nsegs;
            }
        }
        n.ncols += ncol;
        n.nsegs++;
        n.nops++;
        const int lo = r.x0 > 0 ? r.x0 : 0, hi = r.x1 < W - 1 ? r.x1 : W - 1;
        n.reccap += (uint32_t)(hi - lo + 1 > 0 ? hi - lo + 1 : 0);
    }

    // bitmap_render.rs:137-165: is the part `rr` behind the view-space point v?
    FE_HD static bool behind(const RenderRec &rr, V2 v) {
        const float mn = fminf(rr.lsx, rr.lex), mx = fmaxf(rr.lsx, rr.lex);
        if (mn > v.x) return true;
        return mx > v.x && !left_of(v, V2{rr.lsx, rr.lsy}, V2{rr.lex, rr.ley});
    }

    // Every not yet drawn masked mid-texture that -- unless `all` -- lies behind the view-space point v is drawn now, last
    // created first (the reference iterates its reversed part list: mod.rs:124, map_objects.rs:226-232, segs.rs:593-597).
    // Only parts that HAVE something to draw are looked at: for the others (texture "-") being "drawn" changes nothing.
    // dsegs[0 .. nmids) are the mid-textures in creation order (the sprites' headers follow them).
    FE_NOINLINE void draw_parts_behind(bool all, V2 v) {
        for (int c0 = ((int)nmids - 1) / 32 * 32; c0 >= 0 && nmids && n.status == FE_OK; c0 -= 32) {
            uint32_t hit = ballot([&](int l) {
                if (c0 + l >= (int)nmids) return false;
                const RenderRec rr = sc.renders[sc.dseg_part[c0 + l]];
                if (rr.flags & RF_DRAWN) return false;
                return all || behind(rr, v);
            });
            FE_LANES(l) {
                if (hit & (1u << l)) sc.renders[sc.dseg_part[c0 + l]].flags |= RF_DRAWN;
            }
            for (; hit && n.status == FE_OK; hit &= ~(1u << highest(hit))) draw_late(c0 + highest(hit));
        }
        FE_SYNC();
    }

    // The view-independent-state part of one map object (map_objects.rs:40-102): rotation, view transform, field-of-view
    // clip, screen x.  One lane per object.
    struct MoPre {
        float csx, csy, cex, cey, so;
        int32_t sx, ex, pic; // pic: rotation index into Thing::bitmap
        int32_t code;        // 0 not visible, 1 go on, 2 "Clipped line x < -0.01", 3 "Invalid rotation"
    };
    FE_NOINLINE MoPre mo_pre(const Thing &t, float pangle) const {
        const float PI_F = 3.14159265358979323846f;
        MoPre p{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0, 0, 0, 0};
        float ang = pangle - t.angle - PI_F; // map_objects.rs:44-58
        ang += PI_F / 16.0f;
        ang = fmodf(ang, 2.0f * PI_F);
        if (ang < 0.0f) ang += 2.0f * PI_F;
        ang = fmodf(ang, 2.0f * PI_F);
        const uint8_t rotation = as_u8(ang * 8.0f / (2.0f * PI_F));
        if (rotation > 7) {
            p.code = 3;
            return p;
        }
        p.pic = t.rotate ? (int32_t)rotation : 0;
        const Bitmap bm = m.bitmaps[t.bitmap[p.pic]];
        const V2 vpv = rot(sub(V2{t.x, t.y}, ppos), cos_n, sin_n);
        const Seg2 line = {V2{vpv.x - 0.0f, vpv.y - (float)w16(-bm.w) / 2.0f}, V2{vpv.x - 0.0f, vpv.y - (float)bm.w / 2.0f}};
        Seg2 cl;
        float so;
        if (!clip_fov(line, &cl, &so)) return p;
        if (cl.s.x < -0.01f) {
            p.code = 2;
            return p;
        }
        if (t.sector < 0) return p; // "Thing is outside map"
        p.csx = cl.s.x;
        p.csy = cl.s.y;
        p.cex = cl.e.x;
        p.cey = cl.e.y;
        p.so = so;
        const ScreenX sx = project_x(m, cl);
        p.sx = sx.sx;
        p.ex = sx.ex;
        p.code = 1;
        return p;
    }

    // draw_map_objects for one visible object (map_objects.rs:104-214): clip arrays from the parts in front of it, its
    // columns, its BitmapRender.
    FE_NOINLINE void map_object(const Thing &t, const MoPre &pre) {
        const Bitmap bm = m.bitmaps[t.bitmap[pre.pic]];
        const Seg2 cl = {{pre.csx, pre.csy}, {pre.cex, pre.cey}};
        const V2 vpv = rot(sub(V2{t.x, t.y}, ppos), cos_n, sin_n);
        const Sector sec = m.sectors[t.sector];
        const int16_t light = t.full_bright ? (int16_t)255 : sec.light;
        const float ph = pfloor + 41.0f;
        const int16_t z = sec.floor;
        float bh = (float)z - ph, th = (float)z + (float)bm.h - 1.0f - ph;
        bh += (float)t.top_offset[pre.pic] - (float)bm.h;
        th += (float)t.top_offset[pre.pic] - (float)bm.h;
        const ScreenX sx = {pre.sx, pre.ex};
        const ScreenLine bottom = project(m, cl, sx, bh), top = project(m, cl, sx, th);
        const int W = m.W;
        const int16_t H16 = (int16_t)m.H, Hm1 = w16(H16 - 1);
        // The reference builds two [i16; W] clip arrays from EVERY column of every part in front of the sprite
        // (map_objects.rs:104-160) and then reads them at the sprite's own columns [xs, xe) only.  Here a lane owns one sprite
        // column and keeps its two clip values in registers: it visits the parts in front whose columns reach into the
        // sprite's range and looks its own x up in each (a part's columns are sorted by x, usually without gaps).  min / max:
        // the order of the parts does not matter.
        const int xs = (int16_t)bottom.sx, xe = (int16_t)bottom.ex;
        const uint32_t ncol = xe > xs ? (uint32_t)(xe - xs) : 0u;
        if (ncol && (xs < 0 || xe > W)) return fail(FE_PANIC, FED_MO_X);
        if (nallcols + ncol > sc.cap_allcols || nmos + 1 > sc.cap_mos || ndsegs + 1 > sc.cap_dsegs) return fail(FE_HARD, FED_SCRATCH);
        // its columns: x in start.x .. end.x, EXCLUSIVE (map_objects.rs:166; quirk Q6), every one recorded
        const float bd = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
        const float td = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);
        for (int x0 = xs; x0 < xe; x0 += 32) { // 32 sprite columns at a time
            const int x1 = x0 + 32 < xe ? x0 + 32 : xe; // this chunk: [x0, x1)
            PerLane<int32_t> top_clip, bottom_clip;
            FE_LANES(l) {
                top_clip[l] = -1;
                bottom_clip[l] = H16;
            }
            for (int c0 = 0; c0 < (int)nrenders; c0 += 32) {
                uint32_t front = ballot([&](int l) {
                    if (c0 + l >= (int)nrenders) return false;
                    const RenderRec rr = sc.renders[c0 + l];
                    return rr.x1 >= x0 && rr.x0 < x1 && !behind(rr, vpv);
                });
                for (; front; front &= front - 1) {
                    const RenderRec rr = sc.renders[c0 + lowest(front)];
                    const bool dense = (uint32_t)(rr.x1 - rr.x0) + 1u == rr.ncol;
                    FE_LANES(l) {
                        const int x = x0 + l;
                        if (x < x1 && x >= rr.x0 && x <= rr.x1) {
                            uint32_t i = (uint32_t)(x - rr.x0);
                            bool found = true;
                            if (!dense) { // columns hidden behind nearer walls are missing: lower bound on x
                                uint32_t lo = 0, hi = rr.ncol;
                                while (lo < hi) {
                                    const uint32_t mid = (lo + hi) >> 1;
                                    if (sc.allcols[rr.col0 + mid].x < x) lo = mid + 1; else hi = mid;
                                }
                                i = lo;
                                found = lo < rr.ncol && sc.allcols[rr.col0 + lo].x == x;
                            }
                            if (found) {
                                const ColRec c = sc.allcols[rr.col0 + i];
                                if (rr.flags & RF_TWOSIDED) { // map_objects.rs:138-158
                                    if ((rr.flags & RF_DRAW_CEILING) && c.top_y > top_clip[l]) top_clip[l] = c.top_y;
                                    if (c.bottom_y < bottom_clip[l]) bottom_clip[l] = c.bottom_y;
                                } else { // :118-136
                                    if ((rr.flags & RF_EXT_BOTTOM) && c.clipped_top_y < bottom_clip[l]) bottom_clip[l] = c.clipped_top_y;
                                    if ((rr.flags & RF_EXT_TOP) && c.clipped_bottom_y > top_clip[l]) top_clip[l] = c.clipped_bottom_y;
                                }
                            }
                        }
                    }
                }
            }
            FE_LANES(l) {
                const int x = x0 + l;
                if (x < x1) {
                    const int16_t by = as_i16((float)bottom.sy + ((float)(int16_t)x - (float)bottom.sx) * bd);
                    const int16_t ty = as_i16((float)top.sy + ((float)(int16_t)x - (float)top.sx) * td);
                    int16_t ct = ty > (int16_t)top_clip[l] ? ty : (int16_t)top_clip[l], cb = by < (int16_t)bottom_clip[l] ? by : (int16_t)bottom_clip[l];
                    ct = ct < 0 ? (int16_t)0 : ct;
                    cb = Hm1 < cb ? Hm1 : cb;
                    ColRec c;
                    c.x = (int16_t)x;
                    c.clipped_top_y = ct;
                    c.clipped_bottom_y = cb;
                    c.bottom_y = by;
                    c.top_y = ty;
                    sc.allcols[nallcols + (uint32_t)(x - xs)] = c;
                }
            }
        }
        int32_t dseg = -1;
        if (ncol) {
            const Bitmap pb = bm;
            if (pb.slot < 0) return fail(FE_HARD, FED_BITMAP_SLOT);
            dseg = (int32_t)ndsegs;
            FE_LEADER {
                SegRec r;
                r.bitmap_slot = (uint32_t)pb.slot;
                r.light_level = light;
                r.phase = 2;
                r.lsx = cl.s.x;
                r.lsy = cl.s.y;
                r.lex = cl.e.x;
                r.ley = cl.e.y;
                r.start_offset = pre.so;
                r.start_x = bottom.sx;
                r.end_x = bottom.ex;
                r.bottom_height = bh;
                r.top_height = th;
                r.offset_x = 0;
                r.offset_y = 0;
                r.cols_first = nallcols;
                r.n = ncol;
                r.x0 = (int16_t)xs;
                r.x1 = (int16_t)(xe - 1);
                r.tex_base = pb.base;
                r.tex_w = pb.w;
                r.tex_h = pb.h;
                r.tex_opaque = pb.opaque;
                r.pad[0] = r.pad[1] = 0;
                sc.dsegs[ndsegs] = r;
            }
            ndsegs++;
            nallcols += ncol;
        }
        FE_LEADER {
            MoRec mo;
            mo.vx = (cl.s.x + cl.e.x) / 2.0f;
            mo.vy = (cl.s.y + cl.e.y) / 2.0f;
            mo.key = (int32_t)as_i16(cl.s.x);
            mo.dseg = dseg;
            sc.mos[nmos] = mo;
        }
        nmos++;
        FE_SYNC();
    }

    // C: draw_map_objects, map_objects.rs:19-241
    FE_NOINLINE void map_objects(float pangle) {
        for (int c0 = 0; c0 < m.nthings && n.status == FE_OK; c0 += 32) {
            PerLane<float> p_csx, p_csy, p_cex, p_cey, p_so;
            PerLane<int32_t> p_sx, p_ex, p_pic, p_code;
            FE_LANES(l) {
                MoPre p{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0, 0, 0, 0};
                if (c0 + l < m.nthings) p = mo_pre(m.things[c0 + l], pangle);
                p_csx[l] = p.csx;
                p_csy[l] = p.csy;
                p_cex[l] = p.cex;
                p_cey[l] = p.cey;
                p_so[l] = p.so;
                p_sx[l] = p.sx;
                p_ex[l] = p.ex;
                p_pic[l] = p.pic;
                p_code[l] = p.code;
            }
            uint32_t live = ballot([&](int l) { return p_code[l] != 0; });
            for (; live && n.status == FE_OK; live &= live - 1) {
                const int src = lowest(live);
                const MoPre p{from_lane(p_csx, src), from_lane(p_csy, src), from_lane(p_cex, src), from_lane(p_cey, src), from_lane(p_so, src),
                              from_lane(p_sx, src), from_lane(p_ex, src), from_lane(p_pic, src), from_lane(p_code, src)};
                if (p.code == 2) return fail(FE_PANIC, FED_CLIP_X);
                if (p.code == 3) return fail(FE_PANIC, FED_ROTATION);
                map_object(m.things[c0 + src], p);
            }
        }
        if (n.status != FE_OK) return;
        // stable sort on `start.x as i16`, then reverse (map_objects.rs:216-217): descending key, later objects first among
        // equal keys.  Rank by counting.
        FE_LANES(l) {
            for (uint32_t i = (uint32_t)l; i < nmos; i += 32) {
                const int32_t ki = sc.mos[i].key;
                uint32_t rank = 0;
                for (uint32_t j = 0; j < nmos; j++) {
                    const int32_t kj = sc.mos[j].key;
                    rank += (kj > ki || (kj == ki && j > i)) ? 1u : 0u;
                }
                sc.mo_order[rank] = (int32_t)i;
            }
        }
        FE_SYNC();
        // map_objects.rs:219-239: before each object, the masked mid-textures behind it
        for (uint32_t k = 0; k < nmos && n.status == FE_OK; k++) {
            const MoRec mo = sc.mos[sc.mo_order[k]];
            draw_parts_behind(false, V2{mo.vx, mo.vy});
            if (mo.dseg >= 0 && n.status == FE_OK) draw_late(mo.dseg);
        }
    }

    FE_HD bool all_occluded(int xs, int xe) { // every column of [xs, xe] is occluded; 0 <= xs <= xe < W
        const int nslots = (m.W + 1023) >> 10;
        return ballot([&](int l) {
            uint32_t miss = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int s = 0; s < 4; s++) { // (constant slot indices: the masks stay in registers)
                if (s >= nslots) break; // uniform: a 320-column screen has one slot
                // the columns of [xs, xe] inside this lane's word of slot s: bits from .. to, none when from > to
                const int lo = ((s << 5) + l) << 5;
                const int from = xs - lo > 0 ? xs - lo : 0, to = xe - lo < 31 ? xe - lo : 31;
                const uint32_t want = from <= to ? (0xffffffffu >> (31 - to)) & (0xffffffffu << from) : 0u;
                miss |= want & ~occ[s][l];
            }
            return miss == 0u;
        }) == 0xffffffffu;
    }

    // The same question asked by ONE lane for its own seg (the masks are shared memory of the warp): a batch's 32 segs are
    // tested at once before they are taken one by one, so the hidden ones cost nothing serial.  0 <= xs <= xe < W <= 4096.
    FE_HD bool range_occluded(int xs, int xe) {
        for (int w = xs >> 5; w <= (xe >> 5); w++) {
            const int lo = w << 5;
            const int from = xs - lo > 0 ? xs - lo : 0, to = xe - lo < 31 ? xe - lo : 31;
            const uint32_t want = (0xffffffffu >> (31 - to)) & (0xffffffffu << from);
            if (want & ~occ[(w >> 5) & 3].v[w & 31]) return false;
        }
        return true;
    }

    FE_HD SegPre seg_pre(const Seg &sg) const { return seg_pre_of(m, ppos, cos_n, sin_n, sg); }

    // process_seg, segs.rs:353-590 (after seg_pre)
    FE_NOINLINE void seg(const Seg &sg, const SegPre &pre) {
        const Line ld = m.lines[sg.line];
        const int fi = sg.dir ? ld.back : ld.front, bi = sg.dir ? ld.front : ld.back;
        const Side fs = m.sides[fi];
        const Sector fsec = m.sectors[fs.sector];
        const float floor_h = (float)fsec.floor;
        float ceil_h = (float)fsec.ceil;
        bool has_pb = false, has_pt = false;
        float pb = 0.0f, pt = 0.0f;
        Sector bsec = fsec;
        if (bi != -1) {
            bsec = m.sectors[m.sides[bi].sector];
            if (bsec.floor > fsec.floor) {
                has_pb = true;
                pb = (float)bsec.floor;
            }
            if (bsec.ceil < fsec.ceil) {
                has_pt = true;
                pt = (float)bsec.ceil;
            }
        }
        const bool two_sided = (ld.flags & 4) != 0, top_unpeg = (ld.flags & 8) != 0, bottom_unpeg = (ld.flags & 16) != 0;
        if (pre.code == 2) return fail(FE_PANIC, FED_CLIP_X);
        const Seg2 cl = {{pre.csx, pre.csy}, {pre.cex, pre.cey}};
        const float so = pre.so;
        const ScreenX sx = {pre.sx, pre.ex};
        const float ph = pfloor + 41.0f;

        if (fsec.floor_flat < 0 || fsec.ceil_flat < 0) return fail(FE_PANIC, FED_FLAT_MISSING); // Flats::get, flats.rs:92-100
        bool draw_ceiling = true;
        if (bi != -1) { // sky hack, segs.rs:463-477
            if (fsec.ceil_name_has_sky && bsec.ceil_name_has_sky) {
                has_pt = false;
                ceil_h = fminf((float)bsec.ceil, ceil_h);
                draw_ceiling = false;
            }
        }
        if (!two_sided) {
            sidedef_part(cl, sx, so, fs, sg.offset, fsec, floor_h - ph, ceil_h - ph, bottom_unpeg ? as_i32(floor_h - ceil_h) : 0, fs.middle, false,
                         false, false, draw_ceiling, false);
        } else {
            sidedef_part(cl, sx, so, fs, sg.offset, fsec, floor_h - ph, ceil_h - ph, 0, fs.middle, true, false, false, draw_ceiling, false);
            if (n.status != FE_OK) return;
            const float mf = has_pb ? pb : floor_h, mc = has_pt ? pt : ceil_h;
            sidedef_part(cl, sx, so, fs, sg.offset, fsec, mf - ph, mc - ph, 0, fs.middle, false, false, false, draw_ceiling, true);
            if (n.status != FE_OK) return;
            if (has_pb)
                sidedef_part(cl, sx, so, fs, sg.offset, fsec, floor_h - ph, pb - ph, bottom_unpeg ? as_i32(ceil_h - pb) : 0, fs.lower, false, true,
                             false, draw_ceiling, false);
            if (n.status != FE_OK) return;
            if (has_pt)
                sidedef_part(cl, sx, so, fs, sg.offset, fsec, pt - ph, ceil_h - ph, top_unpeg ? 0 : as_i32(pt - ceil_h), fs.upper, false, false, true,
                             draw_ceiling, false);
        }
    }

    // Renderer::render, mod.rs:118-136 (without phase C, the map objects)
    FE_HD void run(const ViewIn &v, const Bases &b) {
        base = b;
        n = Counts{0, 0, 0, 0, 0, 0, FE_OK, FED_NONE, 0, {0, 0, 0}};
        ppos = V2{v.x, v.y};
        cos_n = v.cos_n;
        sin_n = v.sin_n;
        pfloor = 0.0f; // game.rs:144-150, 376-389
        const int s = sector_at(m, ppos);
        if (s >= 0) pfloor = (float)m.sectors[s].floor;
        FE_LANES(l) {
            for (int x = l; x < m.W; x += 32) { // segs.rs:97-99
                sc.hor_ocl[x] = 0;
                sc.floor_ocl[x] = (int16_t)m.H;
                sc.ceil_ocl[x] = (int16_t)-1;
            }
        }
        FE_SYNC();
        open[0] = open[1] = false;
        nrenders = nallcols = ndsegs = nmos = nmids = 0;
        occ_on = m.W <= 4096;
        FE_LANES(l) {
            occ[0][l] = occ[1][l] = occ[2][l] = occ[3][l] = 0u;
        }
        if (EMIT) {
            FE_LEADER { out.views[base.frame] = View{v.x, v.y, pfloor, v.angle, v.cos_a, v.sin_a}; }
        }
        // A: render_node, mod.rs:69-104 -- front subtree, then back subtree (explicit stack instead of the recursion).  The
        // walk itself only decides the ORDER in which the segs are processed (the reference does no occlusion culling in
        // the tree), so it first lists the segs in that order; they are then taken 32 at a time: one lane per seg for the
        // stateless part (seg_pre), the survivors in order through seg().
        int nord = 0;
        int32_t *const order = sc.order;
        const Node *const nodes = m.nodes;
        const SubSector *const ssectors = m.ssectors;
        if (m.node_up) {
            // The order in closed form.  The walk visits, at every node, the child on the player's side first; so a
            // subsector's position in the list is the number of segs under the first-visited siblings along its path to the
            // root.  The 32 lanes evaluate the side test of 32 nodes at a time (one bit each), then every lane climbs from
            // its own subsector to the root adding up -- instead of one lane-uniform pointer chase through every node.
            uint32_t *const side = sc.side;
            for (int i0 = 0; i0 < m.nnodes; i0 += 32) {
                const uint32_t bits = ballot([&](int l) {
                    if (i0 + l >= m.nnodes) return false;
                    const Node nd = nodes[i0 + l];
                    return left_of(ppos, V2{nd.x, nd.y}, V2{nd.x + nd.dx, nd.y + nd.dy});
                });
                FE_LEADER { side[i0 >> 5] = bits; }
            }
            FE_SYNC();
            FE_LANES(l) {
                for (int si = l; si < m.nssectors; si += 32) {
                    int rank = 0;
                    for (int32_t up = m.ss_up[si]; up >= 0;) {
                        const int a = up >> 1;
                        const bool under_left = (up & 1) != 0, is_left = ((side[a >> 5] >> (a & 31)) & 1u) != 0u;
                        const NodeUp nu = m.node_up[a];
                        if (under_left != is_left) rank += under_left ? nu.right : nu.left; // the sibling is visited first
                        up = nu.up;
                    }
                    const SubSector ss = ssectors[si];
                    for (int i = 0; i < ss.count; i++) order[rank + i] = ss.first + i;
                }
            }
            nord = m.nord;
        } else {
        int stack[64];
        int sp = 0;
        stack[sp++] = m.nnodes - 1;
        const int nsegs = m.nsegs;
        while (sp > 0) {
            const int node = stack[--sp];
            if (node < 0) {
                const SubSector ss = ssectors[~node];
                if (nord + ss.count > nsegs) { // subsectors sharing segs: not a map the loaders produce
                    fail(FE_HARD, FED_STACK);
                    break;
                }
                FE_LANES(l) {
                    if (l < ss.count) order[nord + l] = ss.first + l;
                    for (int i = l + 32; i < ss.count; i += 32) order[nord + i] = ss.first + i; // (rare: more than 32 segs)
                }
                nord += ss.count;
                continue;
            }
            const Node nd = nodes[node];
            const V2 a = {nd.x, nd.y}, bb = {nd.x + nd.dx, nd.y + nd.dy};
            const bool is_left = left_of(ppos, a, bb);
            if (sp + 2 > 64) {
                fail(FE_HARD, FED_STACK);
                break;
            }
            stack[sp++] = is_left ? nd.right : nd.left; // visited second
            stack[sp++] = is_left ? nd.left : nd.right; // visited first
        }
        }
        FE_SYNC();
        // The seg numbers are read one batch ahead, and the stateless kernel's code and record of the NEXT batch's segs are
        // requested (prefetch) while this batch's survivors are processed: otherwise every batch starts with three dependent
        // global loads (order -> code -> record).
        PerLane<int32_t> s_cur, s_nxt;
        FE_LANES(l) {
            s_cur[l] = l < nord ? sc.order[l] : -1;
            s_nxt[l] = 32 + l < nord ? sc.order[32 + l] : -1;
        }
        for (int c0 = 0; c0 < nord && n.status == FE_OK; c0 += 32) {
            PerLane<float> p_csx, p_csy, p_cex, p_cey, p_so;
            PerLane<int32_t> p_sx, p_ex, p_code, p_seg;
            FE_LANES(l) {
                SegPre p{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0, 0, 0};
                const int si = s_cur[l], sn = s_nxt[l];
                if (sn >= 0 && sc.pre) {
                    prefetch_l1(sc.pre_code + sn);
                    prefetch_l1(sc.pre + sn);
                }
                if (si >= 0) {
                    if (!sc.pre) p = seg_pre(m.segs[si]);
                    else if (sc.pre_code[si]) p = sc.pre[si]; // the stateless kernel's record: every survivor of the batch fetches its own at once
                }
                s_cur[l] = sn;
                s_nxt[l] = c0 + 64 + l < nord ? sc.order[c0 + 64 + l] : -1;
                p_seg[l] = si >= 0 ? si : 0;
                p_csx[l] = p.csx;
                p_csy[l] = p.csy;
                p_cex[l] = p.cex;
                p_cey[l] = p.cey;
                p_so[l] = p.so;
                p_sx[l] = p.sx;
                p_ex[l] = p.ex;
                p_code[l] = p.code;
            }
            uint32_t live = ballot([&](int l) { return p_code[l] != 0; });
            // what is hidden already (and cannot panic) drops out here, every lane testing its own seg; occlusion only grows, so
            // this is the serial test below asked early -- which remains for what the batch's own earlier segs come to hide
            if (occ_on && live) live &= ~ballot([&](int l) { return p_code[l] == 4 && range_occluded(p_sx[l], p_ex[l]); });
            for (; live && n.status == FE_OK; live &= live - 1) {
                const int src = lowest(live);
                const SegPre p{from_lane(p_csx, src), from_lane(p_csy, src), from_lane(p_cex, src), from_lane(p_cey, src), from_lane(p_so, src),
                               from_lane(p_sx, src), from_lane(p_ex, src), from_lane(p_code, src)};
                if (p.code == 4 && occ_on && all_occluded(p.sx, p.ex)) continue; // nothing of it can be seen, and it cannot panic
                seg(m.segs[from_lane(p_seg, src)], p);
            }
        }
        if (n.status != FE_OK) return;
        if (EMIT && n.nops + n.nplanes > cap.ops) return fail(FE_HARD, FED_CAPACITY);
        // B: mod.rs:106-116 -- the visplanes in push order, after every wall
        if (EMIT) {
            FE_LANES(l) {
                for (uint32_t i = (uint32_t)l; i < n.nplanes; i += 32) out.ops[base.op + n.nops + i] = 0x80000000u | (base.plane + i);
            }
        }
        n.nops += n.nplanes;
        if (m.phases & 4) {
            FE_SYNC();
            nmids = ndsegs;
            if (m.nthings > 0) map_objects(v.angle);       // C: mod.rs:126-133
            if (n.status == FE_OK) draw_parts_behind(true, V2{0.0f, 0.0f}); // D: segs.rs:593-597 -- what is left, last created first
        }
    }
};

// ---- single-pass mode: slabs -> dense lists ------------------------------------------------------------------------
// The emit pass can run WITHOUT a count pass when every view writes into its own fixed-size slab of each array (view v's
// slab of array A starts at v * cap.A).  compact_view then copies a view's lists to
// their final, dense place (the offsets come from an exclusive scan of the counts the emit pass left) and rebases the
// indices they contain, which yields exactly the arrays the two-pass mode writes.  One warp per view.
struct Slabs {
    Out out;  // the slab arrays (views[] is written densely by frame index only in two-pass mode: slab mode indexes it by view)
    Caps cap; // per view
};
FE_HD void compact_view(const Slabs &sl, uint32_t v, const Counts &c, const Bases &b, const Out &dst) {
    const uint32_t s0 = v * sl.cap.segs, c0 = v * sl.cap.cols, o0 = v * sl.cap.ops, p0 = v * sl.cap.planes, r0 = v * sl.cap.parr;
    FE_LEADER { dst.views[b.frame] = sl.out.views[v]; }
    FE_LANES(l) {
        for (uint32_t i = (uint32_t)l; i < c.nops; i += 32) {
            const uint32_t op = sl.out.ops[o0 + i];
            dst.ops[b.op + i] = (op & 0x80000000u) ? (0x80000000u | (b.plane + ((op & 0x7fffffffu) - p0))) : b.seg + (op - s0);
        }
        for (uint32_t i = (uint32_t)l; i < c.nsegs; i += 32) {
            SegRec r = sl.out.segs[s0 + i];
            r.cols_first = b.col + (r.cols_first - c0);
            dst.segs[b.seg + i] = r;
        }
        for (uint32_t i = (uint32_t)l; i < c.ncols; i += 32) dst.cols[b.col + i] = sl.out.cols[c0 + i];
        for (uint32_t i = (uint32_t)l; i < c.nplanes; i += 32) {
            PlaneRec p = sl.out.planes[p0 + i];
            p.arr_first = b.parr + (p.arr_first - r0);
            dst.planes[b.plane + i] = p;
        }
        for (uint32_t i = (uint32_t)l; i < c.nparr; i += 32) dst.parr[b.parr + i] = sl.out.parr[r0 + i];
    }
}

} // namespace fe
} // namespace drr
