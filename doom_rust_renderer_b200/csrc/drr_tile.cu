// drr_tile.cu -- the two sm_100a kernels of the draw path: column binning + span setup, and the tile draw kernel.
//
//   drr_bin_kernel        : one thread per (frame, screen column).  Walks the frame's ops in call order ("column binning":
//                           the emitted lists are per op, the draw kernel wants them per column) and writes, for every
//                           (op, column) that survives clipping, everything of render_vertical_bitmap_line that depends on
//                           the column only (src/renderer/bitmap_render.rs:233-251), the per-visplane constants of
//                           draw_visplane (src/renderer/visplanes.rs:108,112-114) or draw_sky's tx (visplanes.rs:54-66),
//                           decoded all the way into the register values the pixel loops use (64 B per span), so that
//                           the draw kernel spends no instructions on decoding.
//   drr_tile_kernel       : one CTA (8 warps) per (frame, 32 screen columns, band of at most 400 rows).  A warp owns four
//                           adjacent columns; a span belongs to one column, so all of its parameters are uniform over the
//                           8 lanes that work on it, each lane takes TWO rows (y and y + 8) at a time and evaluates them
//                           with Blackwell's packed FP32 instructions (FFMA2 / FMUL2 / FADD2: two independent IEEE f32
//                           results per issue slot).  The per-PIXEL part: wall/sprite ty + texel + diminish_color
//                           (bitmap_render.rs:253-275, 190-208), flat inverse projection (visplanes.rs:103-128), sky
//                           (visplanes.rs:65-77).  Pixels go to a u32 tile in shared memory (layout: tile_offset() in
//                           drr_kernels.h; every store of the draw loops is bank-conflict free); a column's spans are drawn
//                           by ONE lane group in draw order (the kinds that always write overwrite, the kinds with None
//                           texels skip them).  Write-out: a warp reads 16 rows of the tile with four LDS.128 per lane,
//                           packs them to RGB24 (Pixels::set, src/renderer/pixels.rs:22-30) IN PLACE as a row-major 96-byte
//                           x 8-row box per block, and hands the boxes to the TMA unit (cp.async.bulk.tensor, one
//                           instruction per 768 bytes) while the per-frame checksum is accumulated from the packed words.
//                           The palette image (3 KB) arrives by one cp.async.bulk on an mbarrier.
#include "drr_device.cuh"
#include "drr_kernels.h"
#include "drr_math.cuh"
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

namespace drr {

// Decoded span record, 64 bytes, stored as the four 16-byte words a, c, d, b: a flat needs a and c, a wall a, c and d (b only
// when its texture height is not a power of two), so the words a span needs first share a 32-byte sector.  Word layout:
//   all kinds : a.x = y0 | y1 << 16     a.y = kind | flags << 8      a.z = texel index / flat byte offset of the column
//   wall kinds: a.w = K1   d.z = mask   b.y = K2   b.z = magic   b.w = -h      (ty = ((tyr + K1) & mask) + K2, then mod h;
//                                                                              word b is only read when h is not a power of two)
//               c.x = -top_y (f32)   c.y = -(bottom_y - top_y) (f32)   c.z = refined 1/(bottom_y - top_y)   c.w = uy1
//               d.x = bitmap.height as f32 (NaN when bottom_y == top_y)   d.y = light factor   d.w = 2^23 + offset_y (f32, TS_TRUNC)
//   flat      : c.x = wz * vx   c.y = GCFX * wz   c.z = light / 255
static constexpr uint32_t COL_COVERED = 0x80000000u; // ColIdx.n flag: the column's always-writing spans cover every row
enum : uint32_t { TS_POW2 = 1u << 8, TS_BRIGHT = 1u << 9, TS_FASTDIV = 1u << 10, TS_UNIT = 1u << 11, TS_TRUNC = 1u << 12, TS_RUNS = 1u << 13 };
#ifndef DRR_RUN_ROWS
#define DRR_RUN_ROWS 6
#endif
static constexpr int RUN_ROWS = DRR_RUN_ROWS; // consecutive rows a lane takes in the texel-run wall loop (even: rows are evaluated in pairs)

// Decoded record of a wall / sprite column (everything of render_vertical_bitmap_line that depends on the column only)
struct Rec { // a decoded record in registers
    uint4 a, b, c, d;
    uint32_t kind;
};
__device__ __forceinline__ Rec wall_record(const DrawArgs &a, const SegRec &g, const SegConst &kc, int x, int ya, int yb, int top_y, int bottom_y) {
    const uint32_t h = (uint32_t)g.tex_h;
    uint32_t kind = g.tex_opaque ? KIND_WALL : KIND_WALL_HOLES;
    uint4 ra = make_uint4((uint32_t)ya | ((uint32_t)yb << 16), 0u, 0u, 0u), rb = make_uint4(0u, 0u, 0u, 0u), rc = rb, rd = rb;
    const WallColumn wc = wall_column(g, kc, g.tex_w, x);
    if (wc.tx < 0) kind = KIND_NONE; // reference: negative index -> panic
    const uint32_t lp = ilog2_ceil(h); // column-major texel pool: one texture column = 1 << lp consecutive texels
    ra.z = g.tex_base + ((uint32_t)(wc.tx < 0 ? 0 : wc.tx) << lp);
    uint32_t flags = 0;
    if ((h & (h - 1u)) == 0u) {
        // floormod(wrap16(tyr + off_y), 2^k) == (tyr + off_y) & (2^k - 1): the i16 wrap only touches bits >= 16
        flags |= TS_POW2;
        ra.w = (uint32_t)(int)g.offset_y;
        rd.z = h - 1u;
    } else {
        // floormod(v, h) for v in i16 via u = v + M (M = multiple of h >= 32768), q = umulhi(u, magic), r = u - q*h;
        // wrap16(t + off) + M == ((t + off + 32768) & 0xffff) + (M - 32768)
        const uint32_t M = h * ((32768u + h - 1u) / h);
        ra.w = (uint32_t)((int)g.offset_y + 32768);
        rd.z = 0xffffu;
        rb.y = M - 32768u;
        rb.z = (uint32_t)(0x100000000ull / h) + 1u;
        rb.w = 0u - h;
    }
    const int den = bottom_y - top_y;
    const float denF = (float)den;
    rc.x = __float_as_uint(-(float)top_y);
    rc.y = __float_as_uint(-denF);
    rc.z = __float_as_uint(den != 0 ? refined_rcp(denF) : 0.0f);
    rc.w = __float_as_uint(wc.uy1);
    // bottom_y == top_y: ay is NaN or +-inf, (1.0 - ay) * 0.0 is NaN, the sum is NaN and `NaN as i16` is 0
    rd.x = den != 0 ? __float_as_uint((float)h) : 0x7fc00000u;
    rd.y = __float_as_uint(wc.factor);
    if (!(wc.factor <= 1.0f)) flags |= TS_BRIGHT; // light level above 255 or negative depth: channels saturate at 255
    // Are 0 <= sum <= 32767 and 0 <= trunc(sum) + offset_y <= 32767 on every row of the span (sum = ay * uy1 + h)?  The sum is
    // a monotonic function of y (every step of bitmap_render.rs:256-257 is), so its values on the rows ya and yb bound it.
    // Then `sum as i16` is a plain truncation of a non-negative number and `+ offset_y` neither wraps nor goes negative
    // (:259-263 reduce to v mod h), and the pixel loop gets v from ONE round-toward-zero add: the integer part of
    // sum + (2^23 + offset_y) sits in the low mantissa bits (two rows per instruction, no conversion unit).  The two sums are
    // evaluated exactly as the pixel loop evaluates them.
    if (den != 0) {
        const float r = __uint_as_float(rc.z), hF = (float)h, off = (float)(int)g.offset_y;
        const float sa = __fadd_rn(__fmul_rn(fast_div(__fadd_rn((float)ya, -(float)top_y), denF, r), wc.uy1), hF);
        const float sb = __fadd_rn(__fmul_rn(fast_div(__fadd_rn((float)yb, -(float)top_y), denF, r), wc.uy1), hF);
        const float lo = fminf(sa, sb), hi = fmaxf(sa, sb); // (a NaN fails the comparisons below through sa / sb themselves)
        if (sa >= 0.0f && sb >= 0.0f && hi <= 32767.0f && __fadd_rz(lo, off) >= 0.0f && truncf(hi) + off <= 32767.0f) flags |= TS_TRUNC;
        rd.w = __float_as_uint(8388608.0f + off);
        // Magnified walls: when RUN_ROWS consecutive rows advance the texture row by less than one (|uy1| * (RUN_ROWS - 1) <
        // |den|, with a margin for the roundings of the sum), they show at most two different texels, the first row's and the
        // last row's; the texel-run loop then lights two texels per RUN_ROWS pixels instead of one per pixel.
        if ((flags & (TS_TRUNC | TS_BRIGHT)) == TS_TRUNC && yb - ya + 1 >= 2 * RUN_ROWS &&
            fabsf(wc.uy1) * (float)(RUN_ROWS - 1) <= 0.98f * fabsf(denF))
            flags |= TS_RUNS;
    }
    ra.y = kind | flags;
    return Rec{ra, rb, rc, rd, kind};
}

// Decoded record of a visplane column: the per-plane and per-column constants of draw_visplane / draw_sky
__device__ __forceinline__ Rec plane_record(const DrawArgs &a, const PlaneRec &p, const View &vw, int x, int ya, int yb) {
    uint4 ra = make_uint4((uint32_t)ya | ((uint32_t)yb << 16), 0u, 0u, 0u), rc = make_uint4(0u, 0u, 0u, 0u), rd = rc;
    if (p.kind == KIND_FLAT) {
        // visplanes.rs:112  wz = visplane.height as f32 - player.floor_height - PLAYER_EYE_HEIGHT
        const float wz = __fsub_rn(__fsub_rn((float)p.height, vw.floor_height), 41.0f);
        // visplanes.rs:108  vx = (CAMERA_FOCUS_X - x as f32) / ASPECT_RATIO_CORRECTION
        const float vx = __fdiv_rn(__fsub_rn(a.CFX, (float)x), a.ASPECT);
        const float wzvx = __fmul_rn(wz, vx);    // left operand of visplanes.rs:114
        const float gwz = __fmul_rn(a.GCFX, wz); // left operand of visplanes.rs:113
        ra.z = (uint32_t)p.flat_slot * 4096u;
        rc.x = __float_as_uint(wzvx);
        rc.y = __float_as_uint(gwz);
        const float lf = __fdiv_rn((float)p.light_level, 255.0f);
        rc.z = __float_as_uint(lf); // bitmap_render.rs:191
        uint32_t flags = (fast_div_operand_ok(wzvx) && fast_div_operand_ok(gwz)) ? TS_FASTDIV : 0u;
        // Is the light factor <= 1 on every row of the span?  wx = GCFX*wz / (CFY - y) is monotonic in y on either side of the
        // horizon and so is factor = light/255 - (wx as i16)/4096 (every step is a monotonic function), so the rows ya and yb
        // bound it when the span does not touch the horizon row; then the pixel loop needs no saturating path.
        const float vya = __fsub_rn(a.CFY, (float)ya), vyb = __fsub_rn(a.CFY, (float)yb);
        if ((vya > 0.0f) == (vyb > 0.0f) && vya != 0.0f && vyb != 0.0f) {
            const float fa = light_factor(lf, sat_i16(__fdiv_rn(gwz, vya))), fb = light_factor(lf, sat_i16(__fdiv_rn(gwz, vyb)));
            if (fa <= 1.0f && fb <= 1.0f) flags |= TS_UNIT;
        }
        ra.y = KIND_FLAT | flags;
    } else { // sky kinds
        uint32_t kind = (uint32_t)p.kind;
        int tx = sky_tx(vw.angle, x, a.Wf);
        if (tx < 0) { kind = KIND_NONE; tx = 0; }
        ra.z = a.sky_base + ((uint32_t)tx << 7);
        ra.y = kind;
    }
    return Rec{ra, make_uint4(0u, 0u, 0u, 0u), rc, rd, ra.y & 0xffu};
}

// Walk the ops of frame f in call order and visit what each of them draws in screen column x.  EMIT = false only counts;
// EMIT = true writes the decoded records to out[0], out[4], ...  The clipping rules are Pixels::set's (pixels.rs:23: x >= W
// is ignored, rows are clipped to the screen by the callers) and draw_visplane's (visplanes.rs:95-101).  Returns the count.
// s_tab holds, per op of the frame, (x0 | x1 << 16, op word): the x test runs on shared memory, the op's record is only
// loaded on a hit.  Ops beyond the table's capacity are read from global memory.
static constexpr int BIN_THREADS = 384; // upper bound of the CTA size; the launcher picks ceil(W / ceil(W / 192)) rounded up to a warp (the size hardly matters: tools/sweep_env.sh DRR_BIN_THREADS)
static constexpr int BIN_TAB = 512;
#ifndef DRR_BIN_REC
#define DRR_BIN_REC 96
#endif
static constexpr int BIN_REC = DRR_BIN_REC; // ops whose whole record (80-byte SegRec / 16-byte PlaneRec) is staged in shared memory too (A/B: 160 -> walk320 bin 0.191 -> ?)

__device__ __forceinline__ uint2 op_range(const DrawArgs &a, uint32_t op) { // (x0 | x1 << 16, op); an empty op gets x0 > x1
    if (op & 0x80000000u) {
        const PlaneRec p = a.planes[op & 0x7fffffffu];
        return make_uint2((uint32_t)(uint16_t)p.left | ((uint32_t)(uint16_t)p.right << 16), op);
    }
    const SegRec *g = a.segs + op;
    const uint32_t n = g->n;
    const int x0 = n ? g->x0 : 1, x1 = n ? g->x1 : 0;
    return make_uint2((uint32_t)(uint16_t)x0 | ((uint32_t)(uint16_t)x1 << 16), op);
}

// Coverage of a column by the spans that always write (wall / flat / sky without None texels): when their union is the
// whole column, every pixel is written at least once and the tile kernel need not clear the column first.  (In the
// reference's lists a wall and the flat next to it usually share their boundary row, so overlaps are the normal case.)
// Up to COVER_MAX such spans are tracked; more just means "clear it".
static constexpr int COVER_MAX = 6;
struct Cover {
    uint32_t iv[COVER_MAX];
    int n = 0;
    bool overflow = false;
    __device__ __forceinline__ void add(int ya, int yb) {
        if (n < COVER_MAX) {
#pragma unroll
            for (int i = 0; i < COVER_MAX; ++i)
                if (i == n) iv[i] = (uint32_t)ya | ((uint32_t)yb << 16);
            ++n;
        } else {
            overflow = true;
        }
    }
    __device__ __forceinline__ bool covers(int lo, int H) const { // rows [lo, H): sweep, extend the covered prefix [lo, cur) until it stops growing
        if (overflow) return false;
        int cur = lo;
        for (int pass = 0; pass < COVER_MAX && cur < H; ++pass) {
            int reach = cur;
#pragma unroll
            for (int i = 0; i < COVER_MAX; ++i)
                if (i < n && (int)(iv[i] & 0xffffu) <= cur) reach = max(reach, (int)(iv[i] >> 16) + 1);
            if (reach == cur) break;
            cur = reach;
        }
        return cur >= H;
    }
};

// Visit, in call order, the ops of the frame whose x range touches the warp's 32 columns [wx0, wx0 + 31]: 32 table
// entries are tested at a time (one per lane), the survivors are handed to every lane through a ballot.  All 32 lanes of
// the warp must call it together.
template <class F, class P>
__device__ __forceinline__ void for_each_candidate(const DrawArgs &a, uint32_t o0, uint32_t nops, int wx0, const uint2 *s_tab, F visit, P ahead) {
    const int lane = threadIdx.x & 31;
    for (uint32_t base = 0; base < nops; base += 32) {
        const uint32_t k = base + (uint32_t)lane;
        uint2 e = make_uint2(1u, 0u); // empty range
        if (k < nops) e = k < (uint32_t)BIN_TAB ? s_tab[k] : op_range(a, a.ops[o0 + k]);
        const int ex0 = (int)(short)(e.x & 0xffffu), ex1 = (int)(short)(e.x >> 16);
        const bool cand = ex0 <= ex1 && ex1 >= wx0 && ex0 <= wx0 + 31;
        if (cand) ahead(e, k); // the lane that holds a candidate asks for what the warp's 32 columns will read of it, before the serial visits
        uint32_t m = __ballot_sync(0xffffffffu, cand);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            visit(make_uint2(__shfl_sync(0xffffffffu, e.x, src), __shfl_sync(0xffffffffu, e.y, src)), base + (uint32_t)src);
        }
    }
}

// Where the spans of one screen column go: one list per row band (`cap` slots each, band b's list starts at first + b*cap).
struct ColOut {
    uint4 *recs;      // the frame's record slots
    uint32_t first, cap;
    int nlists, band_rows;
    uint32_t n[MAX_LIST_BANDS];
    __device__ __forceinline__ void put(const Rec &r, int ya, int yb) {
        const int b0 = nlists > 1 ? ya / band_rows : 0, b1 = nlists > 1 ? yb / band_rows : 0;
        for (int b = b0; b <= b1; ++b) { // a span that crosses a band boundary is listed in every band it touches
            uint32_t k = 0;
#pragma unroll
            for (int i = 0; i < MAX_LIST_BANDS; ++i)
                if (i == b) k = n[i]++;
            uint4 *out = recs + ((size_t)first + (size_t)b * cap + k) * 4;
            out[0] = r.a;
            out[1] = r.c;
            out[2] = r.d;
            out[3] = r.b;
        }
    }
};

__device__ __forceinline__ void walk_column(const DrawArgs &a, int f, int x, const View &vw, const uint2 *s_tab, const uint4 *s_rec, const SegConst *s_kc, ColOut &out, Cover *cover) {
    const uint32_t o0 = a.frame_op_base[f], nops = a.frame_nops ? a.frame_nops[f] : a.frame_op_base[f + 1] - o0;
    for_each_candidate(a, o0, nops, x & ~31, s_tab, [&](uint2 e, uint32_t k) {
        if (x >= a.W || x < (int)(short)(e.x & 0xffffu) || x > (int)(short)(e.x >> 16)) return; // (Pixels::set ignores x >= W)
        const uint32_t op = e.y;
        if (op & 0x80000000u) {
            const PlaneRec p = k < (uint32_t)BIN_REC ? *reinterpret_cast<const PlaneRec *>(s_rec + 5 * k) : a.planes[op & 0x7fffffffu];
            const uint32_t tb = a.parr[p.arr_first + (uint32_t)(x - p.left)];
            const int t = max((int)(short)(tb & 0xffffu), 0);                // visplanes.rs:61 / :95
            const int b = min((int)(short)(tb >> 16), a.H - 1);              // :62 / :96
            if (p.kind == KIND_FLAT && (int)(short)(b - t) <= 1) return;     // :99-101 (not applied to sky)
            if (t > b) return;
            const Rec r = plane_record(a, p, vw, x, t, b);
            if (r.kind == KIND_FLAT || r.kind == KIND_SKY) cover->add(t, b);
            out.put(r, t, b);
        } else {
            const SegRec *gp = k < (uint32_t)BIN_REC ? reinterpret_cast<const SegRec *>(s_rec + 5 * k) : a.segs + op;
            const uint32_t gn = gp->n, cols_first = gp->cols_first;
            const int gx0 = gp->x0, gx1 = gp->x1;
            // the records' x is strictly increasing: usually x0, x0+1, ... (direct index), otherwise binary search
            uint32_t i = (uint32_t)(x - gx0);
            if ((uint32_t)(gx1 - gx0) + 1u != gn) {
                uint32_t lo = 0, hi = gn;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (a.cols[cols_first + mid].x < x) lo = mid + 1; else hi = mid;
                }
                if (lo >= gn || a.cols[cols_first + lo].x != x) return;
                i = lo;
            }
            const ColRec c = a.cols[cols_first + i];
            const int ya = max((int)c.clipped_top_y, 0), yb = min((int)c.clipped_bottom_y, a.H - 1);
            if (ya > yb) return;
            // the seg's part of the column math: once per op for the staged ones (s_kc), else here
            const Rec r = wall_record(a, *gp, k < (uint32_t)BIN_REC ? s_kc[k] : seg_const(*gp), x, ya, yb, c.top_y, c.bottom_y);
            if (r.kind == KIND_WALL) cover->add(ya, yb);
            out.put(r, ya, yb);
        }
    }, [&](uint2 e, uint32_t k) {
        // what the warp's columns [wx0, wx0 + 31] read of this op first: their (top, bottom) pairs of the visplane (4 bytes per column)
        // or their column records of the seg (10 bytes per column, when the records are dense in x) -- requested into L1 now
        const int wx0 = x & ~31, lo = max(wx0, (int)(short)(e.x & 0xffffu)), hi = min(wx0 + 31, (int)(short)(e.x >> 16));
        const uint32_t op = e.y;
        if (op & 0x80000000u) {
            const PlaneRec p = k < (uint32_t)BIN_REC ? *reinterpret_cast<const PlaneRec *>(s_rec + 5 * k) : a.planes[op & 0x7fffffffu];
            const uint32_t *q = a.parr + p.arr_first + (uint32_t)(lo - p.left);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (hi - lo)));
        } else {
            const SegRec *gp = k < (uint32_t)BIN_REC ? reinterpret_cast<const SegRec *>(s_rec + 5 * k) : a.segs + op;
            if ((uint32_t)(gp->x1 - gp->x0) + 1u == gp->n) {
                const ColRec *q = a.cols + gp->cols_first + (uint32_t)(lo - gp->x0);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (hi - lo) / 2));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (hi - lo)));
            }
        }
    });
}

// drr_bin_kernel: one thread per (frame, screen column).  Counts what the frame's ops draw in the column, reserves that
// many records of the frame's range with one atomic, then writes the records in draw order ("column binning").
#ifndef DRR_BIN_MIN_BLOCKS
#define DRR_BIN_MIN_BLOCKS 3 // 55 registers (tools/sweep_bin_regs.sh: 64 / 55 / 40 / 32 registers -> walk320 bin 0.204 / 0.194 / 0.214 / 0.262 ms)
#endif
__global__ void __launch_bounds__(BIN_THREADS, DRR_BIN_MIN_BLOCKS) drr_bin_kernel(DrawArgs a, int frame0, int bpf) {
    const int nthreads = (int)blockDim.x;
    __shared__ uint2 s_tab[BIN_TAB];
    __shared__ uint4 s_rec[BIN_REC * 5]; // the first BIN_REC ops' records: the walk then depends on one global load (the column record) only
    __shared__ SegConst s_kc[BIN_REC];
    const int f = frame0 + (int)(blockIdx.x / (unsigned)bpf);
    const int x = (int)(blockIdx.x % (unsigned)bpf) * nthreads + (int)threadIdx.x;
    const uint32_t o0 = a.frame_op_base[f], nops = a.frame_nops ? a.frame_nops[f] : a.frame_op_base[f + 1] - o0;
    for (uint32_t k = threadIdx.x; k < min(nops, (uint32_t)BIN_TAB); k += nthreads) s_tab[k] = op_range(a, a.ops[o0 + k]);
    for (uint32_t i = threadIdx.x; i < min(nops, (uint32_t)BIN_REC) * 5; i += nthreads) {
        const uint32_t k = i / 5, part = i % 5, op = a.ops[o0 + k];
        if (op & 0x80000000u) {
            if (part == 0) s_rec[5 * k] = *reinterpret_cast<const uint4 *>(a.planes + (op & 0x7fffffffu));
        } else {
            s_rec[i] = reinterpret_cast<const uint4 *>(a.segs + op)[part];
        }
    }
    __syncthreads();
    // what of the column math depends on the seg only (length, four quotients, light / 255: a square root and five divisions): once per staged op
    for (uint32_t k = threadIdx.x; k < min(nops, (uint32_t)BIN_REC); k += nthreads)
        if (!(s_tab[k].y & 0x80000000u)) s_kc[k] = seg_const(*reinterpret_cast<const SegRec *>(s_rec + 5 * k));
    __syncthreads();
    if ((x & ~31) >= a.W) return; // whole warps only: the candidate walk is a warp-wide operation
    const bool live = x < a.W;
    const View vw = a.views[f];
    // reserve one record per op whose x range contains the column (an upper bound of what survives clipping; the frame's
    // record range is sized by the host from the same bound): this pass touches shared memory only
    uint32_t cap = 0;
    for_each_candidate(a, o0, nops, x & ~31, s_tab, [&](uint2 e, uint32_t) {
        cap += (live && x >= (int)(short)(e.x & 0xffffu) && x <= (int)(short)(e.x >> 16)) ? 1u : 0u;
    }, [](uint2, uint32_t) {});
    // one list of `cap` slots per row band (a single list when the column is not cut into bands, or into too many)
    const int nlists = a.nbands <= MAX_LIST_BANDS ? a.nbands : 1;
    ColOut out;
    out.recs = reinterpret_cast<uint4 *>(a.tparams);
    out.first = a.frame_rec_base[f] * (uint32_t)nlists;
    out.cap = cap;
    out.nlists = nlists;
    out.band_rows = a.band_rows;
#pragma unroll
    for (int i = 0; i < MAX_LIST_BANDS; ++i) out.n[i] = 0;
    if (cap) out.first += atomicAdd(a.frame_cursor + f, cap * (uint32_t)nlists);
    Cover cover;
    // (a dead lane of a partly live warp, x >= W, visits nothing: it only takes part in the ballots)
    walk_column(a, f, x, vw, s_tab, s_rec, s_kc, out, &cover);
    if (live) {
        for (int b = 0; b < nlists; ++b) {
            uint32_t nb = 0;
#pragma unroll
            for (int i = 0; i < MAX_LIST_BANDS; ++i)
                if (i == b) nb = out.n[i];
            const int lo = nlists > 1 ? b * a.band_rows : 0, hi = nlists > 1 ? min(a.H, lo + a.band_rows) : a.H;
            ColIdx ci;
            ci.first = out.first + (uint32_t)b * cap;
            ci.n = nb | (cover.covers(lo, hi) ? COL_COVERED : 0u);
            a.colidx[((size_t)f * nlists + b) * a.W + x] = ci;
        }
    }
}

// sky ty of every screen row (visplanes.rs:68-72: depends on the row only), computed once per context
__global__ void drr_sky_rows_kernel(uint8_t *rows, int H, float Hf) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < H) rows[y] = (uint8_t)sky_ty(y, Hf);
}

// ------------------------------------------------------------------------------------------------------------------
// pixel loops.  A lane group (8 lanes) draws one span: lane li takes the rows ya + li, + 16, ... and, in the same
// iteration, the row 8 below (1 KB further on in the tile); `addr` is the shared address of the lane's first row.
// ------------------------------------------------------------------------------------------------------------------
static constexpr uint32_t ROW8 = 1024, ROW16 = 2048; // tile distance of rows y -> y + 8, y -> y + 16 (tile_offset())

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- Blackwell packed FP32 (FADD2 / FMUL2 / FFMA2: two independent IEEE f32 operations per instruction, same rounding
// as the scalar forms) ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
// ptxas (CUDA 12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 even though both carry an explicit .rn
// and -fmad=false is given (the scalar forms are left alone).  A product that feeds a sum is therefore added with
// fma(p, one, c) where `one` is 1.0f read from the kernel arguments: round(p * 1 + c) == round(p + c), and ptxas cannot
// fold a multiplier it does not know.  tests/test_host.py checks the SASS for it.
__device__ __forceinline__ float2 add2_nofuse(float2 prod, float2 c, float one) { return __ffma2_rn(prod, f2(one), c); }
__device__ __forceinline__ float2 fast_div2(float2 a, float2 negb, float2 r) { // a / b with r = refined 1/b, negb = -b
    const float2 q0 = __fmul2_rn(a, r);
    const float2 rem = __ffma2_rn(negb, q0, a);
    return __ffma2_rn(r, rem, q0);
}
// one pixel's lit colour for 0 <= factor <= 1 (see lit_rgb_unit): (r, g) as a packed pair, b scalar
__device__ __forceinline__ uint32_t lit_rgb_unit_p(float4 p, float factor) {
    const float2 rg = __fadd2_rz(__fmul2_rn(f2(p.x, p.y), f2(factor)), f2(8388608.0f));
    const float b = __fadd_rz(__fmul_rn(p.z, factor), 8388608.0f);
    return __byte_perm(__byte_perm(__float_as_uint(rg.x), __float_as_uint(rg.y), 0x0040), __float_as_uint(b), 0x5410);
}

// palette entry at shared address `addr` as (r, g, b) floats: r and g are bf16 halves of one word, b is an f32
__device__ __forceinline__ float4 pal_fetch(uint32_t addr) {
    uint32_t rg, b;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rg), "=r"(b) : "r"(addr));
    return make_float4(__uint_as_float(rg << 16), __uint_as_float(rg & 0xffff0000u), __uint_as_float(b), 0.0f);
}

// read-only global load of one texel (the pointer's address space is spelled out: its provenance is hidden on purpose, see
// tile_wall_span, and a generic LD would otherwise be emitted)
__device__ __forceinline__ uint32_t ldg_u16(const uint16_t *p) {
    uint16_t v;
    asm("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p) {
    uint32_t v;
    asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

struct TileCtx { // what every pixel loop needs besides its span
    uint32_t pal;      // shared address of the palette image
    uint32_t hole;     // texel pool value of a None texel
    float one;         // 1.0f, opaque to ptxas
    int li;            // lane within its group
};

// bitmap_render.rs:256-265 for the two rows of a lane; yt = (y - top_y) as f32 of both rows.  K = K1 (ra.w).
template <bool POW2>
__device__ __forceinline__ void wall_texels2(const uint4 rb, const uint4 rc, uint32_t K, uint32_t mask, float hF, float2 yt, float one,
                                             const uint16_t *__restrict__ col, uint32_t &t0, uint32_t &t1) {
    const float2 ay = fast_div2(yt, f2(__uint_as_float(rc.y)), f2(__uint_as_float(rc.z)));   // :256
    // :257 with uy0 == 0.0: (1.0 - ay) * 0.0 is +-0.0 for finite ay and h + (+-0.0) == h, so the middle term drops out
    const float2 sum = add2_nofuse(__fmul2_rn(ay, f2(__uint_as_float(rc.w))), f2(hF), one);
    uint32_t u0 = ((uint32_t)sat_i16(sum.x) + K) & mask; // :259
    uint32_t u1 = ((uint32_t)sat_i16(sum.y) + K) & mask;
    if (!POW2) { // :260-263 == floormod (identity checked in tests/: test_wrap_mod_idiom_is_floormod)
        u0 += rb.y;
        u1 += rb.y;
        u0 = __umulhi(u0, rb.z) * rb.w + u0;
        u1 = __umulhi(u1, rb.z) * rb.w + u1;
    }
    t0 = ldg_u16(col + u0);
    t1 = ldg_u16(col + u1);
}
// the same for a TS_TRUNC span: v = trunc(sum) + offset_y is in 0..32767 and comes out of one packed add (magic = 2^23 +
// offset_y, round toward zero); ty = v mod h
template <bool POW2>
__device__ __forceinline__ void wall_texels2_trunc(const uint4 rb, const uint4 rc, float magic, uint32_t mask, float hF, float2 yt, float one,
                                                   const uint16_t *__restrict__ col, uint32_t &t0, uint32_t &t1) {
    const float2 ay = fast_div2(yt, f2(__uint_as_float(rc.y)), f2(__uint_as_float(rc.z)));
    const float2 sum = add2_nofuse(__fmul2_rn(ay, f2(__uint_as_float(rc.w))), f2(hF), one);
    const float2 tr = __fadd2_rz(sum, f2(magic));
    uint32_t u0 = __float_as_uint(tr.x) & mask, u1 = __float_as_uint(tr.y) & mask; // (the mask, <= 0xffff, drops the float's exponent)
    if (!POW2) { // v mod h for 0 <= v < 2^16, h < 2^15: q = (v * (floor(2^32 / h) + 1)) >> 32 is exact
        u0 = __umulhi(u0, rb.z) * rb.w + u0;
        u1 = __umulhi(u1, rb.z) * rb.w + u1;
    }
    t0 = ldg_u16(col + u0);
    t1 = ldg_u16(col + u1);
}

// Two pixels' lit colours for 0 <= factor <= 1: (c as f32 * factor) as u8 per channel is the low byte of c*f + 2^23 rounded
// toward zero (lit_rgb_unit).  The six channels go through three packed multiplies and adds: (g0, b0), (g1, b1), (r0, r1) --
// g is unpacked in place next to b, which the 8-byte palette entry already holds as an f32.
__device__ __forceinline__ void lit_rgb_unit_2(uint2 e0, uint2 e1, float f0, float f1, uint32_t &rgb0, uint32_t &rgb1) {
    const float2 m = f2(8388608.0f);
    const float2 gb0 = __fadd2_rz(__fmul2_rn(f2(__uint_as_float(e0.x & 0xffff0000u), __uint_as_float(e0.y)), f2(f0)), m);
    const float2 gb1 = __fadd2_rz(__fmul2_rn(f2(__uint_as_float(e1.x & 0xffff0000u), __uint_as_float(e1.y)), f2(f1)), m);
    const float2 rr = __fadd2_rz(__fmul2_rn(f2(__uint_as_float(e0.x << 16), __uint_as_float(e1.x << 16)), f2(f0, f1)), m);
    rgb0 = __byte_perm(__byte_perm(__float_as_uint(rr.x), __float_as_uint(gb0.x), 0x0040), __float_as_uint(gb0.y), 0x5410);
    rgb1 = __byte_perm(__byte_perm(__float_as_uint(rr.y), __float_as_uint(gb1.x), 0x0040), __float_as_uint(gb1.y), 0x5410);
}
__device__ __forceinline__ uint2 pal_fetch_raw(uint32_t addr) {
    uint2 e;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
    return e;
}

#ifndef DRR_WALL_UNROLL
#define DRR_WALL_UNROLL 2
#endif
static constexpr int WALL_UNROLL = DRR_WALL_UNROLL; // iterations of the wall loop in flight (A/B: 1, 3, 4)
// the fast wall loop: factor <= 1 and TS_TRUNC (every wall of an ordinary scene)
template <bool HOLES, bool POW2>
__device__ __forceinline__ void tile_wall_span(const TileCtx &t, const uint4 ra, const uint4 rb, const uint4 rc, const uint4 rd, int ya, int yb,
                                               uint32_t addr, const uint16_t *__restrict__ texels) {
    const float hF = __uint_as_float(rd.x), factor = __uint_as_float(rd.y);
    // the span's texture column as ONE 64-bit base, opaque to the compiler: otherwise it re-associates texels + (ra.z + u) and
    // pays a 33-bit add with carry per texel instead of a single IMAD.WIDE.U32 (u * 2 + base)
    const uint16_t *__restrict__ col = texels + ra.z;
    asm("" : "+l"(col));
    const float magic = __uint_as_float(rd.w);
    const int yb8 = yb - TILE_LPG;
    int y = ya + t.li;
    float2 yt = f2(__fadd_rn((float)y, __uint_as_float(rc.x)), __fadd_rn((float)(y + TILE_LPG), __uint_as_float(rc.x)));
#pragma unroll WALL_UNROLL
    for (; y <= yb; y += 2 * TILE_LPG, yt = __fadd2_rn(yt, f2((float)(2 * TILE_LPG))), addr += ROW16) {
        uint32_t t0, t1;
        wall_texels2_trunc<POW2>(rb, rc, magic, rd.z, hF, yt, t.one, col, t0, t1);
        uint32_t rgb0, rgb1;
        lit_rgb_unit_2(pal_fetch_raw(t0), pal_fetch_raw(t1), factor, factor, rgb0, rgb1);
        if (!HOLES || t0 != t.hole) sts_u32(addr, rgb0);
        if (y <= yb8 && (!HOLES || t1 != t.hole)) sts_u32(addr + ROW8, rgb1);
    }
}

// The texel-run wall loop (TS_RUNS: a magnified wall, see wall_record).  A lane takes RUN_ROWS CONSECUTIVE rows; along them the
// texture row v is monotonic and moves by at most one, so they show the texel of the first row (A) or of the last row (B): two
// texel fetches, two palette lookups and one packed lighting per RUN_ROWS pixels; every row still evaluates v (the reference's
// own expression, bit for bit, two rows per packed instruction) to pick A's colour or B's.  The 8 lanes of a group cover
// 8 * RUN_ROWS rows per iteration (RUN_ROWS blocks of the tile).
template <bool HOLES, bool POW2, int K>
__device__ __forceinline__ void tile_wall_span_runs(const TileCtx &t, const uint4 ra, const uint4 rb, const uint4 rc, const uint4 rd, int ya, int yb, int b0,
                                                    uint32_t colx, const uint16_t *__restrict__ texels) {
    static_assert(K % 2 == 0 && K >= 4 && K <= RUN_ROWS, "rows are evaluated in pairs; the span's flag is proven for runs of RUN_ROWS rows");
    const float hF = __uint_as_float(rd.x), factor = __uint_as_float(rd.y), magic = __uint_as_float(rd.w);
    const float2 nden = f2(__uint_as_float(rc.y)), rden = f2(__uint_as_float(rc.z)), uy1 = f2(__uint_as_float(rc.w));
    const uint32_t mask = rd.z;
    const uint16_t *__restrict__ col = texels + ra.z;
    asm("" : "+l"(col));
    int y = ya + K * t.li;
    uint32_t addr[K]; // shared addresses of the lane's rows: K rows anywhere in the block structure, all K blocks further per iteration
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const uint32_t r = (uint32_t)(y + j - b0);
        addr[j] = (colx ^ ((r & 7u) << 4)) + ((r >> 3) << 10);
    }
    float2 yt01 = f2(__fadd_rn((float)y, __uint_as_float(rc.x)), __fadd_rn((float)(y + 1), __uint_as_float(rc.x))); // (y - top_y) as f32 of rows 0, 1
    const float yt_last = __fadd_rn((float)yb, __uint_as_float(rc.x));                                              // ... of the span's last row
    auto vbits = [&](float2 ytp) { // 0x4b000000 + v of two rows (wall_texels2_trunc without the mask)
        const float2 ay = fast_div2(ytp, nden, rden);
        const float2 sum = add2_nofuse(__fmul2_rn(ay, uy1), f2(hF), t.one);
        return __fadd2_rz(sum, f2(magic));
    };
    for (; y <= yb; y += 8 * K, yt01 = __fadd2_rn(yt01, f2((float)(8 * K)))) {
        float2 v[K / 2];
        float2 ytp = yt01;
#pragma unroll
        for (int p = 0; p < K / 2; ++p) {
            // the run's last row is A's partner B; when the run sticks out of the span, B is the span's last row instead (the rows
            // of the run that are inside the span lie between the two)
            if (p == K / 2 - 1) ytp.y = fminf(ytp.y, yt_last);
            v[p] = vbits(ytp);
            ytp = __fadd2_rn(ytp, f2(2.0f));
        }
        const uint32_t bitsA = __float_as_uint(v[0].x);
        uint32_t uA = bitsA & mask, uB = __float_as_uint(v[K / 2 - 1].y) & mask;
        if (!POW2) {
            uA = __umulhi(uA, rb.z) * rb.w + uA;
            uB = __umulhi(uB, rb.z) * rb.w + uB;
        }
        const uint32_t tA = ldg_u16(col + uA), tB = ldg_u16(col + uB);
        uint32_t cA, cB;
        lit_rgb_unit_2(pal_fetch_raw(tA), pal_fetch_raw(tB), factor, factor, cA, cB);
        if (!HOLES || tA != t.hole) sts_u32(addr[0], cA);
#pragma unroll
        for (int j = 1; j < K - 1; ++j) {
            const bool isA = __float_as_uint(j & 1 ? v[j / 2].y : v[j / 2].x) == bitsA;
            if (y + j <= yb && (!HOLES || (isA ? tA : tB) != t.hole)) sts_u32(addr[j], isA ? cA : cB);
        }
        if (y + K - 1 <= yb && (!HOLES || tB != t.hole)) sts_u32(addr[K - 1], cB);
#pragma unroll
        for (int j = 0; j < K; ++j) addr[j] += (uint32_t)K * ROW8;
    }
}

// every other wall span (light level above 255, negative depth, texture rows outside 0..32767, NaN geometry): same
// arithmetic with the saturating conversions spelled out
__device__ __noinline__ void tile_wall_span_any(TileCtx t, uint4 ra, uint4 rb, uint4 rc, uint4 rd, int ya, int yb, uint32_t addr,
                                                const uint16_t *__restrict__ texels) {
    const float hF = __uint_as_float(rd.x), factor = __uint_as_float(rd.y);
    const uint16_t *__restrict__ col = texels + ra.z;
    const bool holes = (ra.y & 0xffu) == KIND_WALL_HOLES, pow2 = (ra.y & TS_POW2) != 0;
    const int yb8 = yb - TILE_LPG;
    int y = ya + t.li;
    float2 yt = f2(__fadd_rn((float)y, __uint_as_float(rc.x)), __fadd_rn((float)(y + TILE_LPG), __uint_as_float(rc.x)));
    for (; y <= yb; y += 2 * TILE_LPG, yt = __fadd2_rn(yt, f2((float)(2 * TILE_LPG))), addr += ROW16) {
        uint32_t t0, t1;
        if (pow2) wall_texels2<true>(rb, rc, ra.w, rd.z, hF, yt, t.one, col, t0, t1);
        else wall_texels2<false>(rb, rc, ra.w, rd.z, hF, yt, t.one, col, t0, t1);
        if (!holes || t0 != t.hole) sts_u32(addr, lit_rgb_any(pal_fetch(t0), factor));
        if (y <= yb8 && (!holes || t1 != t.hole)) sts_u32(addr + ROW8, lit_rgb_any(pal_fetch(t1), factor));
    }
}

// visplanes.rs:103-128 for one pixel of a flat span, every division IEEE
__device__ __forceinline__ uint32_t flat_pixel_slow(float vy, float gwz, float wzvx, float lf, float cos_a, float sin_a, int px16, int py16,
                                                    const uint8_t *__restrict__ flat, uint32_t pal) {
    const float wx = __fdiv_rn(gwz, vy), wy = __fdiv_rn(wzvx, vy);
    const float rx = __fsub_rn(__fmul_rn(wx, cos_a), __fmul_rn(wy, sin_a)); // vertexes.rs:20-25
    const float ry = __fadd_rn(__fmul_rn(wy, cos_a), __fmul_rn(wx, sin_a));
    const uint32_t tx = (uint32_t)(sat_i16(rx) + px16); // i16 wrap does not reach the low 6 bits
    const uint32_t ty = (uint32_t)(sat_i16(ry) + py16);
    const uint32_t texel = flat[((ty << 6) & 0xfc0u) | (tx & 63u)];
    return lit_rgb_any(pal_fetch(pal + texel * PAL_ENTRY), light_factor(lf, sat_i16(wx)));
}

struct FlatView { // per-frame constants of the flat loops
    float CFY, cos_a, sin_a;
    int px16, py16;
};

// the fast flat loop: TS_FASTDIV and TS_UNIT (the span does not touch the horizon row, factor <= 1 on every row).  UNROLL = 2
// in the 48-register build of the tall tiles (measured: -1.2 % at 1280x800, +2.7 % in the 40-register build at 320x200).
template <int UNROLL>
__device__ __forceinline__ void tile_flat_span(const TileCtx &t, const FlatView &v, const uint4 ra, const uint4 rc, int ya, int yb, uint32_t addr,
                                               const uint8_t *__restrict__ flats) {
    const float wzvx = __uint_as_float(rc.x), gwz = __uint_as_float(rc.y), lf = __uint_as_float(rc.z);
    const uint8_t *__restrict__ flat = flats + ra.z; // ONE 64-bit base, opaque to the compiler (see tile_wall_span)
    asm("" : "+l"(flat));
    const int yb8 = yb - TILE_LPG, py6 = v.py16 << 6;
    int y = ya + t.li;
    // rows y and y + 8 of this lane together (visplanes.rs:109-128, twice)
    float2 vy = f2(__fsub_rn(v.CFY, (float)y), __fsub_rn(v.CFY, (float)(y + TILE_LPG)));
#pragma unroll UNROLL
    for (; y <= yb; y += 2 * TILE_LPG, vy = __fadd2_rn(vy, f2((float)(-2 * TILE_LPG))), addr += ROW16) {
        float2 r0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.x) : "f"(vy.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.y) : "f"(vy.y));
        const float2 nvy = f2(-vy.x, -vy.y);
        const float2 r = __ffma2_rn(r0, __ffma2_rn(nvy, r0, f2(1.0f)), r0); // refined_rcp, twice
        const float2 wx = fast_div2(f2(gwz), nvy, r);                       // :113
        const float2 wy = fast_div2(f2(wzvx), nvy, r);                      // :114
        const float2 rx = add2_nofuse(__fmul2_rn(wx, f2(v.cos_a)), __fmul2_rn(wy, f2(-v.sin_a)), t.one); // wx*cos - wy*sin (vertexes.rs:20-25)
        const float2 ry = add2_nofuse(__fmul2_rn(wy, f2(v.cos_a)), __fmul2_rn(wx, f2(v.sin_a)), t.one);
        // :119-122  tx = (rx as i16 + px) & 63, ty = (ry as i16 + py) & 63 (the i16 wrap does not reach the low 6 bits); the flat
        // is row-major: offset = ty * 64 + tx
        const uint32_t o0 = ((uint32_t)(sat_i16(ry.x) * 64 + py6) & 0xfc0u) | ((uint32_t)(sat_i16(rx.x) + v.px16) & 63u);
        const uint32_t o1 = ((uint32_t)(sat_i16(ry.y) * 64 + py6) & 0xfc0u) | ((uint32_t)(sat_i16(rx.y) + v.px16) & 63u);
        const uint32_t t0 = ldg_u8(flat + o0), t1 = ldg_u8(flat + o1);
        // diminish_color :191-201: light/255 - dist * (1/4096), clamped below at 0
        const float2 dist = f2((float)sat_i16(wx.x), (float)sat_i16(wx.y));
        // (dist is an integer below 2^15, so dist / 4096 is exact and one fused multiply-add rounds like the reference's two steps)
        float2 fac = __ffma2_rn(dist, f2(-0.000244140625f), f2(lf));
        // `if factor < 0.0 { factor = 0.0 }`: fmaxf turns -0.0 into +0.0, which changes nothing once multiplied and cast to u8
        fac.x = fmaxf(fac.x, 0.0f);
        fac.y = fmaxf(fac.y, 0.0f);
        uint32_t rgb0, rgb1;
        lit_rgb_unit_2(pal_fetch_raw(t.pal + t0 * PAL_ENTRY), pal_fetch_raw(t.pal + t1 * PAL_ENTRY), fac.x, fac.y, rgb0, rgb1);
        sts_u32(addr, rgb0);
        if (y <= yb8) sts_u32(addr + ROW8, rgb1);
    }
}

// every other flat span: one row per lane and iteration, the IEEE division where the hoisted reciprocal is not proven
// (operands outside 2^-60..2^60, the horizon row vy == 0), the saturating colour path where the factor may exceed 1
__device__ __noinline__ void tile_flat_span_any(TileCtx t, FlatView v, uint4 ra, uint4 rc, int ya, int yb, uint32_t addr, const uint8_t *__restrict__ flats) {
    const float wzvx = __uint_as_float(rc.x), gwz = __uint_as_float(rc.y), lf = __uint_as_float(rc.z);
    const uint8_t *__restrict__ flat = flats + ra.z;
    const bool fast = (ra.y & TS_FASTDIV) != 0;
    int y = ya + t.li;
    float vy = __fsub_rn(v.CFY, (float)y);
    for (; y <= yb; y += TILE_LPG, vy = __fadd_rn(vy, (float)-TILE_LPG), addr += ROW8) {
        uint32_t rgb;
        if (fast && vy != 0.0f) {
            const float r = refined_rcp(vy);
            const float wx = fast_div(gwz, vy, r), wy = fast_div(wzvx, vy, r);
            const float rx = __fsub_rn(__fmul_rn(wx, v.cos_a), __fmul_rn(wy, v.sin_a));
            const float ry = __fadd_rn(__fmul_rn(wy, v.cos_a), __fmul_rn(wx, v.sin_a));
            const uint32_t tx = (uint32_t)(sat_i16(rx) + v.px16), ty = (uint32_t)(sat_i16(ry) + v.py16);
            const uint32_t texel = flat[((ty << 6) & 0xfc0u) | (tx & 63u)];
            rgb = lit_rgb_any(pal_fetch(t.pal + texel * PAL_ENTRY), light_factor(lf, sat_i16(wx)));
        } else {
            rgb = flat_pixel_slow(vy, gwz, wzvx, lf, v.cos_a, v.sin_a, v.px16, v.py16, flat, t.pal);
        }
        sts_u32(addr, rgb);
    }
}

__device__ __forceinline__ void tile_sky_span(const TileCtx &t, const uint4 ra, int ya, int yb, uint32_t addr, const uint8_t *__restrict__ sky_rows,
                                              const uint16_t *__restrict__ texels) {
    const bool holes = (ra.y & 0xffu) == KIND_SKY_HOLES;
    // packed 0x00BBGGRR word of the entry whose 8-byte slot is at shared address e: pal + SM_PAL_PACKED + (e - pal) / 2
    const uint32_t packed = t.pal - (t.pal >> 1) + SM_PAL_PACKED;
    for (int y = ya + t.li; y <= yb; y += TILE_LPG, addr += ROW8) {
        const uint32_t texel = texels[ra.z + sky_rows[y]]; // column-major sky: base + tx*128 + ty
        if (holes && texel == t.hole) continue;
        sts_u32(addr, lds_u32(packed + (texel >> 1))); // no lighting (visplanes.rs:74-77)
    }
}

// ---- asynchronous copies (mbarrier, bulk copy, TMA store) ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy (16-byte aligned, a multiple of 16 bytes), completion counted on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global: one box of the framebuffer tensor (96 bytes x 8 rows) from a dense row-major image of it in shared memory
__device__ __forceinline__ void tma_store_box(const CUtensorMap *map, uint32_t src, int x_bytes, int row, int slot) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(x_bytes), "r"(row), "r"(slot) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
// Timing experiments (results are WRONG with any bit set): a build with -DDRR_DBG_KNOBS reads DrawArgs::dbg (env DRR_DBG):
// 1 skip clearing, 2 skip the TMA stores, 4 skip the checksum, 8 skip the write-out, 16 skip drawing.  The shipped library
// has none of these tests.
#ifdef DRR_DBG_KNOBS
#define DBG(bit) ((a.dbg & (bit)) != 0)
#else
#define DBG(bit) false
#endif
// MINB = resident CTAs per SM the register budget is set for (6: tiles of up to ~224 rows; 4: up to 400 rows)
// FAST = W % 32 == 0 and every band a multiple of 8 rows: TMA write-out with the checksum fused; otherwise a bytewise
//        write-out and a separate checksum pass
#ifndef DRR_TILE_MINB_SMALL
#define DRR_TILE_MINB_SMALL 6 // resident CTAs per SM the small-tile build is compiled for (A/B: 5 = 48 registers)
#endif
template <int MINB, bool FAST>
__global__ void __launch_bounds__(TILE_THREADS, MINB) drr_tile_kernel(const __grid_constant__ DrawArgs a, const __grid_constant__ CUtensorMap fbmap, int frame0) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    // Every shared address below is derived from the PROBED base in the kernel arguments (a constant-bank operand), not from
    // cvta(s_dyn) (which ptxas re-derives from SR_CgaCtaId wherever it runs out of registers); the two must agree, since the
    // texel pool holds absolute shared addresses of palette entries.
    if ((uint32_t)__cvta_generic_to_shared(s_dyn) != a.pal_base) __trap();
    const uint32_t pal = a.pal_base + SM_PAL, bar = a.pal_base + SM_BAR, tile = a.pal_base + SM_TILE;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); // (the shuffle tells ptxas it is warp-uniform)
    const int grp = lane >> 3, li = lane & 7;
    const uint16_t *__restrict__ texels = a.texels;
    const uint8_t *__restrict__ flats = a.flats;

    // grid: x = column group, y = row band, z = frame of this launch
    const int g = (int)blockIdx.x, band = (int)blockIdx.y;
    const int f = frame0 + (int)blockIdx.z;
    const int b0 = band * a.band_rows, b1 = min(a.H, b0 + a.band_rows) - 1;
    const int nlists = a.nbands <= MAX_LIST_BANDS ? a.nbands : 1, lband = a.nbands <= MAX_LIST_BANDS ? band : 0; // the bin kernel's list of this band
    // the palette image: one bulk copy, completion on an mbarrier
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_proxy_async(); // the initialised barrier is visible to the copy engine
        mbar_expect_tx(bar, SM_PAL_BYTES);
        bulk_load(pal, a.pal_image, SM_PAL_BYTES, bar);
    }
    // everything else the CTA needs from global memory is requested before the first barrier, so that the round trips overlap
    const View vw = a.views[f];
    const uint32_t slot = a.frame_slot[f];
    const int c = warp * 4 + grp, x = g * TILE_COLS + c; // a warp owns four adjacent columns, a lane group one of them
    ColIdx ci;
    ci.first = 0; ci.n = 0;
    // (frame * (lists per frame * W) as ONE 32 x 32 -> 64 bit multiply-add: nlists * W < 2^18, lband * W + x too)
    if (x < a.W) ci = a.colidx[(size_t)(uint32_t)f * (uint32_t)(nlists * a.W) + (uint32_t)(lband * a.W + x)];
    int left = DBG(16) ? 0 : (int)(ci.n & ~COL_COVERED); // spans of the column still to draw
    uint32_t rec = ci.first;                             // record of the next one
    const uint4 *__restrict__ P = reinterpret_cast<const uint4 *>(a.tparams);
    // The next span's first word is fetched while the current one is drawn: into registers where the budget allows (MINB 4),
    // otherwise only into L1 (four registers kept across the pixel loops would spill at 40)
    constexpr bool HEAD_IN_REGS = MINB <= 5;
#ifdef DRR_RUNS_ALWAYS
    constexpr bool RUNS = true;
#elif defined(DRR_RUNS_NEVER)
    constexpr bool RUNS = false;
#else
    constexpr bool RUNS = MINB <= 5; // the texel-run wall loop: in the 48 / 56 register builds of the tall tiles (spills at 40)
#endif
    // rows per run: RUN_ROWS, or 4 in the 40-register build (no spill there; a span flagged for runs of RUN_ROWS rows qualifies for shorter ones)
    constexpr int RUN_K = MINB >= 6 ? 4 : RUN_ROWS;
    // ... and the record REC_AHEAD spans further on is requested into L1 (64 bytes each, consecutive per column).  Measured (tile
    // kernel ms, 1 / 2 / 3 / 4 ahead): walk1280 0.8124 / 0.8049 / 0.8064 / 0.8073, things640 1.878 / 1.834 / 1.839 / 1.842, stress1920
    // 41.99 / 40.77 / 41.04 / 40.96 -- but walk320 (no head in registers, 40 registers) 0.5282 / 0.5304 / 0.5321 / 0.5333.
#ifdef DRR_REC_AHEAD
    constexpr int REC_AHEAD = DRR_REC_AHEAD;
#else
    constexpr int REC_AHEAD = HEAD_IN_REGS ? 2 : 1;
#endif
    uint4 ra_next = make_uint4(0u, 0u, 0u, 0u);
    if (left > 0) {
        if (HEAD_IN_REGS) ra_next = P[(size_t)rec * 4]; // the first span's head: in flight while the palette arrives
        else asm volatile("prefetch.global.L1 [%0];" ::"l"(P + (size_t)rec * 4));
#pragma unroll
        for (int k = 1; k < REC_AHEAD; ++k)
            if (left > k) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(P + ((size_t)rec + k) * 4));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(P + ((size_t)rec + k) * 4 + 2));
            }
    }
    __syncthreads(); // the barrier's initialisation is visible to every thread
    mbar_wait(bar, 0);

    // ---- draw: each lane group walks its column's span list in draw order ("last writer wins, transparent texels do not
    // write"): the kinds that always write simply overwrite, the kinds with None texels skip them
    {
        TileCtx t;
        t.pal = pal;
        t.hole = a.pal_base + SM_PAL + TEXEL_NONE_INDEX * PAL_ENTRY;
        t.one = a.one;
        t.li = li;
        FlatView fv;
        fv.CFY = a.CFY;
        fv.cos_a = vw.cos_a;
        fv.sin_a = vw.sin_a;
        fv.px16 = sat_i16(vw.pos_x); // visplanes.rs:119-120 `player.position.x as i16`
        fv.py16 = sat_i16(vw.pos_y);
        // shared address of row r (of the band) of this lane group's column, tile_offset(c, r): the tile base is 128-byte
        // aligned, so the row's slot (bits 4..6, swizzled with bit 2 of the quad = warp) goes in with an exclusive or
        uint32_t colx = (tile + ((uint32_t)warp << 7) + ((uint32_t)grp << 2)) ^ (((uint32_t)warp & 4u) << 4);
#ifndef DRR_NO_PIN
        // (opaque to ptxas: otherwise the 40-register build re-derives it from SR_TID.X -- ~18 instructions -- in front of every span
        // instead of keeping one register)
        if (MINB >= 6) asm volatile("" : "+r"(colx), "+r"(t.li));
#endif
        auto row_addr = [&](int r) { return (colx ^ (((uint32_t)r & 7u) << 4)) + (((uint32_t)r >> 3) << 10); };
        // uncovered pixels are (0,0,0) like the reference's zero-initialised Pixels::new (pixels.rs:10-14); a column whose
        // always-writing spans cover every row (the normal case, flagged by the bin kernel) needs no clearing
        if (!(ci.n & COL_COVERED) && !DBG(1)) {
            uint32_t addr = row_addr(li);
            for (int r = li; r <= b1 - b0; r += TILE_LPG, addr += ROW8) sts_u32(addr, 0u);
        }
        while (__any_sync(0xffffffffu, left > 0)) {
            __syncwarp(); // a span may overwrite what another lane of the group stored for an earlier span of the column
            if (left > 0) {
                const uint4 *__restrict__ R = P + (size_t)rec * 4;
                const uint4 ra = HEAD_IN_REGS ? ra_next : R[0];
                ++rec;
                if (--left > 0) { // the next span's first word is fetched while this one is drawn (word c arrives in L1 with it) ...
                    if (HEAD_IN_REGS) ra_next = R[4];
                    if (REC_AHEAD == 1) {
                        if (!HEAD_IN_REGS) asm volatile("prefetch.global.L1 [%0];" ::"l"(R + 4));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(R + 6)); // ... and so is the sector with words d and b
                    } else if (left >= REC_AHEAD) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(R + 4 * REC_AHEAD));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(R + 4 * REC_AHEAD + 2));
                    }
                }
                const int ya = max((int)(ra.x & 0xffff), b0), yb = min((int)(ra.x >> 16), b1);
                const uint32_t kind = ra.y & 0xffu;
                if (ya <= yb && kind != KIND_NONE) {
                    const uint32_t addr = row_addr(ya + li - b0);
                    if (kind == KIND_FLAT) {
                        // a span clipped to a band stays inside the rows its flags were computed for
                        if ((ra.y & (TS_UNIT | TS_FASTDIV)) == (TS_UNIT | TS_FASTDIV)) tile_flat_span<(MINB <= 5 ? 2 : 1)>(t, fv, ra, R[1], ya, yb, addr, flats);
                        else tile_flat_span_any(t, fv, ra, R[1], ya, yb, addr, flats);
                    } else if (kind <= KIND_WALL_HOLES) {
                        const uint4 rc = R[1], rd = R[2];
                        if ((ra.y & (TS_TRUNC | TS_BRIGHT)) != TS_TRUNC) tile_wall_span_any(t, ra, R[3], rc, rd, ya, yb, addr, texels);
                        else if (RUNS && (ra.y & TS_RUNS)) {
                            if (kind == KIND_WALL) {
                                if (ra.y & TS_POW2) tile_wall_span_runs<false, true, RUN_K>(t, ra, ra, rc, rd, ya, yb, b0, colx, texels);
                                else tile_wall_span_runs<false, false, RUN_K>(t, ra, R[3], rc, rd, ya, yb, b0, colx, texels);
                            } else {
                                if (ra.y & TS_POW2) tile_wall_span_runs<true, true, RUN_K>(t, ra, ra, rc, rd, ya, yb, b0, colx, texels);
                                else tile_wall_span_runs<true, false, RUN_K>(t, ra, R[3], rc, rd, ya, yb, b0, colx, texels);
                            }
                        }
                        else if (kind == KIND_WALL) {
                            if (ra.y & TS_POW2) tile_wall_span<false, true>(t, ra, ra, rc, rd, ya, yb, addr, texels);
                            else tile_wall_span<false, false>(t, ra, R[3], rc, rd, ya, yb, addr, texels);
                        } else {
                            if (ra.y & TS_POW2) tile_wall_span<true, true>(t, ra, ra, rc, rd, ya, yb, addr, texels);
                            else tile_wall_span<true, false>(t, ra, R[3], rc, rd, ya, yb, addr, texels);
                        }
                    } else if (kind == KIND_SKY || kind == KIND_SKY_HOLES) {
                        tile_sky_span(t, ra, ya, yb, addr, a.sky_rows, texels);
                    }
                }
            }
        }
    }
    __syncthreads(); // every span of the tile is in before the write-out

    // ---- write-out: Pixels::set (pixels.rs:22-30), RGB24 at 3*(y*W + x)
    const int nrows = b1 - b0 + 1;
    if (FAST) {
        // A warp step = 16 rows of the tile (two blocks) = 16 x 96 bytes of the framebuffer.  Lane (rhi, h, rlo) reads the
        // 16 pixels of half h (quads 4h .. 4h+3) of row rr = 4*rhi + rlo: four LDS.128, conflict-free (a quarter warp = 4 rows
        // x 2 halves, see tile_offset()); packs them into 48 bytes; and, once every lane of the warp has read, stores them at
        // rr*96 + h*48 of the SAME two blocks (conflict-free too), which then hold two dense 96-byte x 8-row boxes for the TMA.
        const int rlo = lane & 3, h = (lane >> 2) & 1, rr = (lane >> 3) * 4 + rlo;
        const int nblk = nrows >> 3;
        const uint32_t src_off = ((uint32_t)(rr >> 3) << 10) + ((uint32_t)h << 9) + ((((uint32_t)rr & 7u) ^ ((uint32_t)h << 2)) << 4);
        const uint32_t dst_off = (uint32_t)rr * 96u + (uint32_t)h * 48u;
        // checksum (drr_device.cuh): this lane's twelve words of a step are one 48-byte group of the frame, number
        // row * (row pitch / 48) + 2 * column group + half; its weight in step s is k0 + s * kstep
        const uint32_t gpr = (uint32_t)a.W / 16u; // groups per framebuffer row
        uint32_t k0 = ((uint32_t)(b0 + 16 * warp + rr) * gpr + 2u * (uint32_t)g + (uint32_t)h + 1u) * CK_C;
        const uint32_t kstep = 16u * (TILE_THREADS / 32) * gpr * CK_C;
        uint64_t acc = 0;
        if (!DBG(8))
        for (int s = warp; 2 * s < nblk; s += TILE_THREADS / 32, k0 += kstep) {
            const uint32_t region = tile + ((uint32_t)s << 11);
            const bool act = 2 * s + (rr >> 3) < nblk; // (the last step of a band of 8 * odd rows has one block only)
            uint4 v0, v1, v2;
            if (act) {
                const uint4 q0 = lds_u128(region + src_off), q1 = lds_u128(region + src_off + 128u);
                const uint4 q2 = lds_u128(region + src_off + 256u), q3 = lds_u128(region + src_off + 384u);
                // 4 pixels (0x??BBGGRR each) -> 3 words of the byte stream R G B R | G B R G | B R G B
                v0.x = __byte_perm(q0.x, q0.y, 0x4210); v0.y = __byte_perm(q0.y, q0.z, 0x5421); v0.z = __byte_perm(q0.z, q0.w, 0x6542);
                v0.w = __byte_perm(q1.x, q1.y, 0x4210); v1.x = __byte_perm(q1.y, q1.z, 0x5421); v1.y = __byte_perm(q1.z, q1.w, 0x6542);
                v1.z = __byte_perm(q2.x, q2.y, 0x4210); v1.w = __byte_perm(q2.y, q2.z, 0x5421); v2.x = __byte_perm(q2.z, q2.w, 0x6542);
                v2.y = __byte_perm(q3.x, q3.y, 0x4210); v2.z = __byte_perm(q3.y, q3.z, 0x5421); v2.w = __byte_perm(q3.z, q3.w, 0x6542);
            }
            __syncwarp(); // every lane has read its pixels: the two blocks may be overwritten
            if (act) {
                sts_u128(region + dst_off, v0);
                sts_u128(region + dst_off + 16u, v1);
                sts_u128(region + dst_off + 32u, v2);
            }
            fence_proxy_async(); // this lane's stores are visible to the copy engine ...
            __syncwarp();        // ... and so are everybody else's, before lane 0 hands the boxes over
            if (lane == 0 && !DBG(2)) {
                tma_store_box(&fbmap, region, g * (TILE_COLS * 3), b0 + 16 * s, (int)slot);
                if (2 * s + 1 < nblk) tma_store_box(&fbmap, region + 768u, g * (TILE_COLS * 3), b0 + 16 * s + 8, (int)slot);
                bulk_commit();
            }
            if (act && !DBG(4)) {
                uint32_t sum = v0.x * checksum_word_weight(0);
                sum += v0.y * checksum_word_weight(1); sum += v0.z * checksum_word_weight(2); sum += v0.w * checksum_word_weight(3);
                sum += v1.x * checksum_word_weight(4); sum += v1.y * checksum_word_weight(5); sum += v1.z * checksum_word_weight(6); sum += v1.w * checksum_word_weight(7);
                sum += v2.x * checksum_word_weight(8); sum += v2.y * checksum_word_weight(9); sum += v2.z * checksum_word_weight(10); sum += v2.w * checksum_word_weight(11);
                acc += (uint64_t)sum * k0;
            }
        }
        // the warp's sum mod 2^64 through three REDUX instead of five rounds of 64-bit shuffles: 22-bit limbs, so that 32 of them add
        // up without a carry out of 32 bits
        {
            const uint32_t l0 = (uint32_t)acc & 0x3fffffu, l1 = (uint32_t)(acc >> 22) & 0x3fffffu, l2 = (uint32_t)(acc >> 44);
            const uint32_t s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1), s2 = __reduce_add_sync(0xffffffffu, l2);
            acc = (uint64_t)s0 + ((uint64_t)s1 << 22) + ((uint64_t)s2 << 44);
        }
        if (lane == 0) {
            if (acc) atomicAdd(reinterpret_cast<unsigned long long *>(a.crc + slot), (unsigned long long)acc);
            bulk_wait_read_all(); // the copy engine has read this warp's boxes: the CTA's shared memory may be handed on
        }
    } else {
        const size_t pitch = (size_t)a.W * 3;
        uint8_t *base = a.frames + (size_t)slot * a.frame_stride + (size_t)g * (TILE_COLS * 3);
        for (int i = threadIdx.x; i < TILE_COLS * nrows; i += TILE_THREADS) { // generic widths: bytewise; the checksum is a separate pass
            const int cc = i % TILE_COLS, r = i / TILE_COLS;
            if (g * TILE_COLS + cc >= a.W) continue;
            const uint32_t rgb = lds_u32(tile + tile_offset(cc, r));
            uint8_t *p = base + (size_t)(b0 + r) * pitch + cc * 3;
            p[0] = (uint8_t)rgb;
            p[1] = (uint8_t)(rgb >> 8);
            p[2] = (uint8_t)(rgb >> 16);
        }
    }
}

// shared window address of a CTA's dynamic shared memory (no static shared memory, like the tile kernel)
__global__ void drr_shared_base_kernel(uint32_t *out) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    if (threadIdx.x == 0) *out = (uint32_t)__cvta_generic_to_shared(s_dyn);
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
cudaError_t probe_shared_base(uint32_t *base) {
    uint32_t *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    drr_shared_base_kernel<<<1, 32, SM_TILE + 1024>>>(d);
    e = cudaMemcpy(base, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e;
}

cudaError_t launch_bin(const DrawArgs &a, int frame0, int nframes, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.frame_cursor + frame0, 0, sizeof(uint32_t) * (size_t)nframes, st);
    if (e != cudaSuccess) return e;
    const int maxt = 192;                                           // (the CTA size hardly matters: 128 .. 384 measured)
    const int bpf = (a.W + maxt - 1) / maxt;                        // CTAs per frame
    const int threads = ((a.W + bpf - 1) / bpf + 31) / 32 * 32;     // columns per CTA, whole warps
    const long long blocks = (long long)nframes * bpf;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    drr_bin_kernel<<<(unsigned)blocks, threads, 0, st>>>(a, frame0, bpf);
    return cudaGetLastError();
}

cudaError_t launch_sky_rows(uint8_t *rows, int H, cudaStream_t st) {
    drr_sky_rows_kernel<<<(H + 255) / 256, 256, 0, st>>>(rows, H, (float)(uint32_t)H);
    return cudaGetLastError();
}

void tile_bands(int H, int max_rows, int *nbands, int *band_rows) {
    // rows per band: equal bands, a multiple of 8 rows when H is, so that every band is whole 8-row blocks (the TMA write-out's
    // box).  Up to 800 rows: bands of at most 400 rows (a tile is then 51 KB: four CTAs per SM; five CTAs of 272 rows were
    // measured equal at 1280x800, and cutting a 400-row screen in two costs 6 %).  Taller screens: at most 328 rows, so that
    // five CTAs fit (1920x1200: 0.798 -> 0.724 ms per 128 frames of the stress map).  The bin kernel writes one span list
    // per (column, band); max_rows > 0 overrides (A/B runs: DRR_TILE_MAX_ROWS).
    if (max_rows <= 0) max_rows = H > 800 ? 328 : 400;
    max_rows = std::max(8, std::min(max_rows, 400)) / 8 * 8;
    *nbands = (H + max_rows - 1) / max_rows;
    *band_rows = (H + *nbands - 1) / *nbands;
    if (H % 8 == 0) *band_rows = (*band_rows + 7) / 8 * 8;
}

void tile_config(int W, int H, int *tc, int *lpg) {
    *tc = TILE_COLS;
    *lpg = TILE_LPG;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember which (device, kernel) pairs have it
static cudaError_t allow_big_tiles(const void *kernel) {
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, kernel})) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) done.insert({dev, kernel});
    return e;
}

template <int MINB, bool FAST>
static cudaError_t launch_tile_t(const DrawArgs &a, const CUtensorMap &map, int frame0, int nframes, size_t dyn, cudaStream_t st, int *launches) {
    const int gpf = (a.W + TILE_COLS - 1) / TILE_COLS;
    cudaError_t e = allow_big_tiles(reinterpret_cast<const void *>(&drr_tile_kernel<MINB, FAST>));
    if (e != cudaSuccess) return e;
    for (int f0 = 0; f0 < nframes; f0 += 65535) { // gridDim.z limit
        const dim3 grid((unsigned)gpf, (unsigned)a.nbands, (unsigned)std::min(65535, nframes - f0));
        drr_tile_kernel<MINB, FAST><<<grid, TILE_THREADS, dyn, st>>>(a, map, frame0 + f0);
        ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_tile(const DrawArgs &a, const CUtensorMap *fbmap, int frame0, int nframes, cudaStream_t st, int *launches) {
    *launches = 0;
    if (nframes <= 0) return cudaSuccess;
    const int gpf = (a.W + TILE_COLS - 1) / TILE_COLS;
    if (a.nbands > 65535) return cudaErrorInvalidConfiguration; // gridDim.y limit
    const size_t dyn = SM_TILE + (size_t)((a.band_rows + 7) / 8) * 1024; // whole 8-row blocks
    const bool fast = fbmap && a.W % TILE_COLS == 0 && a.H % 8 == 0 && a.band_rows % 8 == 0;
    // resident CTAs per SM by tile size: six (40 registers) up to 36 KB, five (48 registers) up to 44 KB, else four (56 registers,
    // with the texel-run wall loop)
    const int minb = dyn <= 36 * 1024 ? DRR_TILE_MINB_SMALL : dyn <= 44 * 1024 ? 5 : 4;
    cudaError_t e;
    if (fast) {
        e = minb == DRR_TILE_MINB_SMALL ? launch_tile_t<DRR_TILE_MINB_SMALL, true>(a, *fbmap, frame0, nframes, dyn, st, launches)
            : minb == 5 ? launch_tile_t<5, true>(a, *fbmap, frame0, nframes, dyn, st, launches) : launch_tile_t<4, true>(a, *fbmap, frame0, nframes, dyn, st, launches);
    } else {
        static const CUtensorMap none = {};
        e = minb == DRR_TILE_MINB_SMALL ? launch_tile_t<DRR_TILE_MINB_SMALL, false>(a, none, frame0, nframes, dyn, st, launches)
            : minb == 5 ? launch_tile_t<5, false>(a, none, frame0, nframes, dyn, st, launches) : launch_tile_t<4, false>(a, none, frame0, nframes, dyn, st, launches);
        if (e == cudaSuccess) e = launch_checksum_pass(a, frame0, nframes, st, launches);
    }
    return e;
}

} // namespace drr
