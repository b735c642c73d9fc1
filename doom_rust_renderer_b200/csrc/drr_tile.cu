// drr_tile.cu -- the two sm_100a kernels of the draw path: column binning + span setup, and the tile draw kernel.
//
//   drr_bin_kernel        : one thread per (frame, screen column).  Walks the frame's ops in call order ("column binning":
//                           the emitted lists are per op, the draw kernel wants them per column) and writes, for every
//                           (op, column) that survives clipping, everything of render_vertical_bitmap_line that depends on
//                           the column only (src/renderer/bitmap_render.rs:233-251), the per-visplane constants of
//                           draw_visplane (src/renderer/visplanes.rs:108,112-114) or draw_sky's tx (visplanes.rs:54-66),
//                           decoded all the way into the register values the pixel loops use (64 B per span), so that
//                           the draw kernel spends no instructions on decoding.
//   drr_tile_kernel       : one CTA per (frame, TC screen columns, band of rows; the band is the whole column for
//                           H <= 800).  A span belongs to one screen column, so all of its parameters are uniform over
//                           the lanes that work on it: a group of LPG lanes (the whole warp for tall screens, 16 or 8 lanes
//                           for short ones, so that short spans still fill the warp) takes a span, each lane takes TWO rows
//                           (y and y + LPG) at a time and evaluates them with Blackwell's packed FP32 instructions
//                           (FFMA2 / FMUL2 / FADD2: two independent IEEE f32 results per issue slot).  The per-PIXEL part:
//                           wall/sprite ty + texel + diminish_color (bitmap_render.rs:253-275, 190-208), flat inverse
//                           projection (visplanes.rs:103-128), sky (visplanes.rs:65-77).  Pixels go to a column-major u32
//                           tile in shared memory (consecutive lanes -> consecutive words: conflict-free); a column's
//                           spans are drawn by ONE lane group in draw order (the kinds that always write overwrite, the
//                           kinds with None texels skip them), and finally the tile is written out row by row as 16-byte
//                           vectors of the row-major RGB24 framebuffer (Pixels::set, src/renderer/pixels.rs:22-30) while
//                           the per-frame checksum is accumulated.  Column sets are handed to warps dynamically (shared
//                           counter), so a warp that drew short columns takes more of them.
#include "drr_device.cuh"
#include "drr_kernels.h"
#include "drr_math.cuh"
#include <algorithm>

namespace drr {

// Decoded span record, 64 bytes.  Word layout (a.x .. d.w):
//   all kinds : a.x = y0 | y1 << 16     a.y = kind | flags << 8      a.z = texel index / flat byte offset of the column
//   wall kinds: a.w = K1   b.x = mask   b.y = K2   b.z = magic   b.w = -h      (ty = ((tyr + K1) & mask) + K2, then mod h)
//               c.x = -top_y (f32)   c.y = -(bottom_y - top_y) (f32)   c.z = refined 1/(bottom_y - top_y)   c.w = uy1
//               d.x = bitmap.height as f32 (NaN when bottom_y == top_y)   d.y = light factor
//   flat      : c.x = wz * vx   c.y = GCFX * wz   c.z = light / 255
static constexpr uint32_t COL_COVERED = 0x80000000u; // ColIdx.n flag: the column's always-writing spans cover every row
enum : uint32_t { TS_POW2 = 1u << 8, TS_BRIGHT = 1u << 9, TS_FASTDIV = 1u << 10, TS_UNIT = 1u << 11 };

// Decoded record of a wall / sprite column (everything of render_vertical_bitmap_line that depends on the column only)
struct Rec { // a decoded record in registers
    uint4 a, b, c, d;
    uint32_t kind;
};
__device__ __forceinline__ Rec wall_record(const DrawArgs &a, const SegRec &g, int x, int ya, int yb, int top_y, int bottom_y) {
    const uint32_t h = (uint32_t)g.tex_h;
    uint32_t kind = g.tex_opaque ? KIND_WALL : KIND_WALL_HOLES;
    uint4 ra = make_uint4((uint32_t)ya | ((uint32_t)yb << 16), 0u, 0u, 0u), rb = make_uint4(0u, 0u, 0u, 0u), rc = rb, rd = rb;
    const WallColumn wc = wall_column(g, g.tex_w, x);
    if (wc.tx < 0) kind = KIND_NONE; // reference: negative index -> panic
    const uint32_t lp = ilog2_ceil(h); // column-major texel pool: one texture column = 1 << lp consecutive texels
    ra.z = g.tex_base + ((uint32_t)(wc.tx < 0 ? 0 : wc.tx) << lp);
    uint32_t flags = 0;
    if ((h & (h - 1u)) == 0u) {
        // floormod(wrap16(tyr + off_y), 2^k) == (tyr + off_y) & (2^k - 1): the i16 wrap only touches bits >= 16
        flags |= TS_POW2;
        ra.w = (uint32_t)(int)g.offset_y;
        rb.x = h - 1u;
    } else {
        // floormod(v, h) for v in i16 via u = v + M (M = multiple of h >= 32768), q = umulhi(u, magic), r = u - q*h;
        // wrap16(t + off) + M == ((t + off + 32768) & 0xffff) + (M - 32768)
        const uint32_t M = h * ((32768u + h - 1u) / h);
        ra.w = (uint32_t)((int)g.offset_y + 32768);
        rb.x = 0xffffu;
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
    ra.y = kind | flags;
    return Rec{ra, rb, rc, rd, kind};
}

// Decoded record of a visplane column: the per-plane and per-column constants of draw_visplane / draw_sky
__device__ __forceinline__ Rec plane_record(const DrawArgs &a, const PlaneRec &p, const View &vw, int x, int ya, int yb) {
    uint4 ra = make_uint4((uint32_t)ya | ((uint32_t)yb << 16), 0u, 0u, 0u), rc = make_uint4(0u, 0u, 0u, 0u);
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
    return Rec{ra, make_uint4(0u, 0u, 0u, 0u), rc, make_uint4(0u, 0u, 0u, 0u), ra.y & 0xffu};
}

// Walk the ops of frame f in call order and visit what each of them draws in screen column x.  EMIT = false only counts;
// EMIT = true writes the decoded records to out[0], out[4], ...  The clipping rules are Pixels::set's (pixels.rs:23: x >= W
// is ignored, rows are clipped to the screen by the callers) and draw_visplane's (visplanes.rs:95-101).  Returns the count.
// s_tab holds, per op of the frame, (x0 | x1 << 16, op word): the x test runs on shared memory, the op's record is only
// loaded on a hit.  Ops beyond the table's capacity are read from global memory.
static constexpr int BIN_THREADS = 384; // upper bound of the CTA size; the launcher picks ceil(W / ceil(W / 192)) rounded up to a warp (the size hardly matters: tools/sweep_env.sh DRR_BIN_THREADS)
static constexpr int BIN_TAB = 512;
static constexpr int BIN_REC = 96; // ops whose whole record (80-byte SegRec / 16-byte PlaneRec) is staged in shared memory too

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
template <class F>
__device__ __forceinline__ void for_each_candidate(const DrawArgs &a, uint32_t o0, uint32_t nops, int wx0, const uint2 *s_tab, F visit) {
    const int lane = threadIdx.x & 31;
    for (uint32_t base = 0; base < nops; base += 32) {
        const uint32_t k = base + (uint32_t)lane;
        uint2 e = make_uint2(1u, 0u); // empty range
        if (k < nops) e = k < (uint32_t)BIN_TAB ? s_tab[k] : op_range(a, a.ops[o0 + k]);
        const int ex0 = (int)(short)(e.x & 0xffffu), ex1 = (int)(short)(e.x >> 16);
        uint32_t m = __ballot_sync(0xffffffffu, ex0 <= ex1 && ex1 >= wx0 && ex0 <= wx0 + 31);
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
            out[1] = r.b;
            out[2] = r.c;
            out[3] = r.d;
        }
    }
};

__device__ __forceinline__ void walk_column(const DrawArgs &a, int f, int x, const View &vw, const uint2 *s_tab, const uint4 *s_rec, ColOut &out, Cover *cover) {
    const uint32_t o0 = a.frame_op_base[f], nops = a.frame_op_base[f + 1] - o0;
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
            const Rec r = wall_record(a, *gp, x, ya, yb, c.top_y, c.bottom_y);
            if (r.kind == KIND_WALL) cover->add(ya, yb);
            out.put(r, ya, yb);
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
    const int f = frame0 + (int)(blockIdx.x / (unsigned)bpf);
    const int x = (int)(blockIdx.x % (unsigned)bpf) * nthreads + (int)threadIdx.x;
    const uint32_t o0 = a.frame_op_base[f], nops = a.frame_op_base[f + 1] - o0;
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
    if ((x & ~31) >= a.W) return; // whole warps only: the candidate walk is a warp-wide operation
    const bool live = x < a.W;
    const View vw = a.views[f];
    // reserve one record per op whose x range contains the column (an upper bound of what survives clipping; the frame's
    // record range is sized by the host from the same bound): this pass touches shared memory only
    uint32_t cap = 0;
    for_each_candidate(a, o0, nops, x & ~31, s_tab, [&](uint2 e, uint32_t) {
        cap += (live && x >= (int)(short)(e.x & 0xffffu) && x <= (int)(short)(e.x >> 16)) ? 1u : 0u;
    });
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
    walk_column(a, f, x, vw, s_tab, s_rec, out, &cover);
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
// pixel loops
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

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

static constexpr uint32_t TEXEL_HOLE = TEXEL_NONE; // texel pool values are palette byte offsets (index * PAL_ENTRY); entry 256 = None

// palette entry at shared address `addr` as (r, g, b) floats
__device__ __forceinline__ float4 pal_fetch(uint32_t addr) {
#ifdef DRR_PAL8
    uint32_t rg, b;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rg), "=r"(b) : "r"(addr));
    return make_float4(__uint_as_float(rg << 16), __uint_as_float(rg & 0xffff0000u), __uint_as_float(b), 0.0f);
#else
    return lds_f4(addr);
#endif
}

// read-only global load of one texel (the pointer's address space is spelled out: its provenance is hidden on purpose, see
// tile_wall_span, and a generic LD would otherwise be emitted)
__device__ __forceinline__ uint32_t ldg_u16(const uint16_t *p) {
    uint16_t v;
    asm("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}

// bitmap_render.rs:256-265 for the two rows of a lane; yt = (y - top_y) as f32 of both rows
template <bool POW2>
__device__ __forceinline__ void wall_texels2(const DrawArgs &a, const uint4 ra, const uint4 rb, const uint4 rc, float hF, float2 yt, float one,
                                             const uint16_t *__restrict__ texels, uint32_t &t0, uint32_t &t1) {
    const float2 ay = fast_div2(yt, f2(__uint_as_float(rc.y)), f2(__uint_as_float(rc.z)));   // :256
    // :257 with uy0 == 0.0: (1.0 - ay) * 0.0 is +-0.0 for finite ay and h + (+-0.0) == h, so the middle term drops out
    const float2 sum = add2_nofuse(__fmul2_rn(ay, f2(__uint_as_float(rc.w))), f2(hF), one);
    uint32_t u0 = ((uint32_t)sat_i16(sum.x) + ra.w) & rb.x; // :259
    uint32_t u1 = ((uint32_t)sat_i16(sum.y) + ra.w) & rb.x;
    if (!POW2) { // :260-263 == floormod (identity checked in tests/: test_wrap_mod_idiom_is_floormod)
        u0 += rb.y;
        u1 += rb.y;
        u0 = __umulhi(u0, rb.z) * rb.w + u0;
        u1 = __umulhi(u1, rb.z) * rb.w + u1;
    }
#ifdef DRR_TEXFETCH
    t0 = tex1Dfetch<unsigned short>(a.tex_texels, (int)(ra.z + u0)); // TEX pipe: beside the LSU/L1 data pipe the kernel is bound by
    t1 = tex1Dfetch<unsigned short>(a.tex_texels, (int)(ra.z + u1));
#else
    t0 = ldg_u16(texels + u0); // `texels` already points at the span's texture column
    t1 = ldg_u16(texels + u1);
#endif
}

template <int LPG, bool HOLES, bool POW2>
__device__ __forceinline__ void tile_wall_span(const DrawArgs &a, const uint4 ra, const uint4 rb, const uint4 rc, const uint4 rd, int ya, int yb, int b0, int li,
                                               uint32_t col_addr, const uint16_t *__restrict__ texels, uint32_t pal_addr, float one) {
    const float hF = __uint_as_float(rd.x), factor = __uint_as_float(rd.y);
    // the span's texture column as ONE 64-bit base, opaque to the compiler: otherwise it re-associates texels + (ra.z + u) and
    // pays a 33-bit add with carry per texel instead of a single IMAD.WIDE.U32 (u * 2 + base)
    const uint16_t *__restrict__ col = texels + ra.z;
    asm("" : "+l"(col));
    int y = ya + li;
    uint32_t addr = col_addr + 4u * (uint32_t)(y - b0);
    float2 yt = f2(__fadd_rn((float)y, __uint_as_float(rc.x)), __fadd_rn((float)(y + LPG), __uint_as_float(rc.x)));
    if (!(ra.y & TS_BRIGHT)) {
#pragma unroll 2
        for (; y <= yb; y += 2 * LPG, yt = __fadd2_rn(yt, f2((float)(2 * LPG))), addr += 8u * LPG) {
            uint32_t t0, t1;
            wall_texels2<POW2>(a, ra, rb, rc, hF, yt, one, col, t0, t1);
            const uint32_t rgb0 = lit_rgb_unit_p(pal_fetch(pal_addr + t0), factor), rgb1 = lit_rgb_unit_p(pal_fetch(pal_addr + t1), factor);
            if (!HOLES || t0 != TEXEL_HOLE) sts_u32(addr, rgb0);
            if (y + LPG <= yb && (!HOLES || t1 != TEXEL_HOLE)) sts_u32(addr + 4u * LPG, rgb1);
        }
    } else {
        for (; y <= yb; y += 2 * LPG, yt = __fadd2_rn(yt, f2((float)(2 * LPG))), addr += 8u * LPG) {
            uint32_t t0, t1;
            wall_texels2<POW2>(a, ra, rb, rc, hF, yt, one, col, t0, t1);
            if (!HOLES || t0 != TEXEL_HOLE) sts_u32(addr, lit_rgb(pal_fetch(pal_addr + t0), factor));
            if (y + LPG <= yb && (!HOLES || t1 != TEXEL_HOLE)) sts_u32(addr + 4u * LPG, lit_rgb(pal_fetch(pal_addr + t1), factor));
        }
    }
}

// visplanes.rs:103-128 for one pixel of a flat span, every division IEEE
__device__ __forceinline__ uint32_t flat_pixel_slow(float vy, float gwz, float wzvx, float lf, float cos_a, float sin_a, int px16, int py16,
                                                    const uint8_t *__restrict__ flat, uint32_t pal_addr) {
    const float wx = __fdiv_rn(gwz, vy), wy = __fdiv_rn(wzvx, vy);
    const float rx = __fsub_rn(__fmul_rn(wx, cos_a), __fmul_rn(wy, sin_a)); // vertexes.rs:20-25
    const float ry = __fadd_rn(__fmul_rn(wy, cos_a), __fmul_rn(wx, sin_a));
    const uint32_t tx = (uint32_t)(sat_i16(rx) + px16); // i16 wrap does not reach the low 6 bits
    const uint32_t ty = (uint32_t)(sat_i16(ry) + py16);
    const uint32_t texel = flat[((ty << 6) & 0xfc0u) | (tx & 63u)];
    return lit_rgb_any(pal_fetch(pal_addr + texel * PAL_ENTRY), light_factor(lf, sat_i16(wx)));
}

template <int LPG, bool UNIT>
__device__ __forceinline__ void tile_flat_span(const uint4 ra, const uint4 rc, int ya, int yb, int b0, int li, uint32_t col_addr, float CFY,
                                               float cos_a, float sin_a, int px16, int py16, const uint8_t *__restrict__ flats,
                                               uint32_t pal_addr, float one, cudaTextureObject_t tex_flats) {
    const float wzvx = __uint_as_float(rc.x), gwz = __uint_as_float(rc.y), lf = __uint_as_float(rc.z);
    const uint8_t *__restrict__ flat = flats + ra.z;
    int y = ya + li;
    uint32_t addr = col_addr + 4u * (uint32_t)(y - b0);
    if (ra.y & TS_FASTDIV) {
        // rows y and y + LPG of this lane together (visplanes.rs:109-128, twice).  The row with vy == 0 (y == H/2 for even H)
        // divides by zero: the loop leaves garbage there (no fault), it is redone below with the IEEE division.
        float2 vy = f2(__fsub_rn(CFY, (float)y), __fsub_rn(CFY, (float)(y + LPG)));
#ifdef DRR_FLAT_UNROLL2
#pragma unroll 2
#endif
        for (; y <= yb; y += 2 * LPG, vy = __fadd2_rn(vy, f2((float)(-2 * LPG))), addr += 8u * LPG) {
            float2 r0;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.x) : "f"(vy.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.y) : "f"(vy.y));
            const float2 nvy = f2(-vy.x, -vy.y);
            const float2 r = __ffma2_rn(r0, __ffma2_rn(nvy, r0, f2(1.0f)), r0); // refined_rcp, twice
            const float2 wx = fast_div2(f2(gwz), nvy, r);                       // :113
            const float2 wy = fast_div2(f2(wzvx), nvy, r);                      // :114
            const float2 rx = add2_nofuse(__fmul2_rn(wx, f2(cos_a)), __fmul2_rn(wy, f2(-sin_a)), one); // wx*cos - wy*sin (vertexes.rs:20-25)
            const float2 ry = add2_nofuse(__fmul2_rn(wy, f2(cos_a)), __fmul2_rn(wx, f2(sin_a)), one);
            const uint32_t tx0 = (uint32_t)(sat_i16(rx.x) + px16), ty0 = (uint32_t)(sat_i16(ry.x) + py16);
            const uint32_t tx1 = (uint32_t)(sat_i16(rx.y) + px16), ty1 = (uint32_t)(sat_i16(ry.y) + py16);
#ifdef DRR_TEXFETCH
            const uint32_t t0 = tex1Dfetch<unsigned char>(tex_flats, (int)(ra.z + (((ty0 << 6) & 0xfc0u) | (tx0 & 63u))));
            const uint32_t t1 = tex1Dfetch<unsigned char>(tex_flats, (int)(ra.z + (((ty1 << 6) & 0xfc0u) | (tx1 & 63u))));
#else
            const uint32_t t0 = flat[((ty0 << 6) & 0xfc0u) | (tx0 & 63u)];
            const uint32_t t1 = flat[((ty1 << 6) & 0xfc0u) | (tx1 & 63u)];
#endif
            // diminish_color :191-201: light/255 - dist * (1/4096), clamped below at 0
            const float2 dist = f2((float)sat_i16(wx.x), (float)sat_i16(wx.y));
            float2 fac = add2_nofuse(__fmul2_rn(dist, f2(-0.000244140625f)), f2(lf), one);
            // `if factor < 0.0 { factor = 0.0 }`: fmaxf turns -0.0 into +0.0, which changes nothing once multiplied and cast to u8
            fac.x = fmaxf(fac.x, 0.0f);
            fac.y = fmaxf(fac.y, 0.0f);
            const float4 p0 = pal_fetch(pal_addr + t0 * PAL_ENTRY), p1 = pal_fetch(pal_addr + t1 * PAL_ENTRY);
            uint32_t rgb0, rgb1;
            if (UNIT || (fac.x <= 1.0f && fac.y <= 1.0f)) {
                rgb0 = lit_rgb_unit_p(p0, fac.x);
                rgb1 = lit_rgb_unit_p(p1, fac.y);
            } else {
                rgb0 = lit_rgb_any(p0, fac.x);
                rgb1 = lit_rgb_any(p1, fac.y);
            }
            sts_u32(addr, rgb0);
            if (y + LPG <= yb) sts_u32(addr + 4u * LPG, rgb1);
        }
        const int ym = (int)CFY; // exact integer when H is even
        if (!UNIT && (float)ym == CFY && ym >= ya && ym <= yb && li == ((ym - ya) % LPG)) // (a UNIT span does not touch that row)
            sts_u32(col_addr + 4u * (uint32_t)(ym - b0), flat_pixel_slow(0.0f, gwz, wzvx, lf, cos_a, sin_a, px16, py16, flat, pal_addr));
    } else {
        float vy = __fsub_rn(CFY, (float)y);
        for (; y <= yb; y += LPG, vy -= (float)LPG, addr += 4u * LPG)
            sts_u32(addr, flat_pixel_slow(vy, gwz, wzvx, lf, cos_a, sin_a, px16, py16, flat, pal_addr));
    }
}

template <int LPG, bool HOLES>
__device__ __forceinline__ void tile_sky_span(const uint4 ra, int ya, int yb, int b0, int li, uint32_t col_addr,
                                              const uint8_t *__restrict__ sky_rows, const uint16_t *__restrict__ texels, uint32_t pal_addr) {
    uint32_t addr = col_addr + 4u * (uint32_t)(ya + li - b0);
    for (int y = ya + li; y <= yb; y += LPG, addr += 4u * LPG) {
        const uint32_t texel = texels[ra.z + sky_rows[y]]; // column-major sky: base + tx*128 + ty
        if (HOLES && texel == TEXEL_HOLE) continue;
#ifdef DRR_PAL8
        sts_u32(addr, lds_u32(pal_addr + 257u * 8u + (texel >> 1))); // packed RGB table behind the 8-byte entries; no lighting (visplanes.rs:74-77)
#else
        sts_u32(addr, lds_u32(pal_addr + texel + 12u)); // no lighting (visplanes.rs:74-77)
#endif
    }
}

// PRMT selector that assembles an output word starting at channel `ph` of pixel a: bytes a[ph..2] then b[0..]
__device__ __forceinline__ uint32_t wsel(int ph) { return ph == 0 ? 0x4210u : ph == 1 ? 0x5421u : 0x6542u; }

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
// TC  = screen columns per tile (16: 48-byte row segments, 3 lanes x 10 rows per write-out step; 32: 96 bytes, 6 lanes x 5 rows)
// LPG = lanes per span (32, 16 or 8); a warp works on 32 / LPG adjacent columns at once
// RP  = tile column pitch in words: >= rows, RP % 32 == 2 (TC 16) or 1 (TC 32) so that the write-out reads are conflict-free
template <int TC, int LPG, int NT, int NSPLIT, bool FAST_STORE>
__global__ void __launch_bounds__(NT, TC == 32 ? TILE_MIN_BLOCKS_SHORT : TILE_MIN_BLOCKS) drr_tile_kernel(const __grid_constant__ DrawArgs a, int frame0, int band_rows, int nbands, int RP) {
    extern __shared__ uint32_t s_tile[]; // [TC columns][RP] u32 pixels (0x00BBGGRR)
#ifdef DRR_PAL8
    __shared__ __align__(16) uint32_t s_pal[257 * 2 + 257]; // 257 x (bf16 r | bf16 g << 16, f32 b), then 257 packed 0x00BBGGRR
#else
    __shared__ float4 s_pal[257];        // entry 256 backs the None texel (its colour is never stored)
#endif
    __shared__ int s_next;
    constexpr int G = 32 / LPG;          // columns per warp step
    constexpr int NSETS = TC / G;
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gpf = (a.W + TC - 1) / TC;
    const int grp = lane / LPG, li = lane % LPG;
    const uint32_t pal_addr = (uint32_t)__cvta_generic_to_shared(s_pal);
    const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(s_tile);
    const uint16_t *__restrict__ texels = a.texels;
    const uint8_t *__restrict__ flats = a.flats;

    // grid: x = column group (times the row band when the column is cut into bands), y = frame of this launch
    int g = (int)blockIdx.x, band = 0;
    if (nbands > 1) {
        g = (int)(blockIdx.x / (unsigned)nbands);
        band = (int)blockIdx.x - g * nbands;
    }
    const int f = frame0 + (int)blockIdx.y;
    const int b0 = band * band_rows, b1 = min(a.H, b0 + band_rows) - 1;
    const int nlists = nbands <= MAX_LIST_BANDS ? nbands : 1, lband = nbands <= MAX_LIST_BANDS ? band : 0; // the bin kernel's list of this band
    // everything the CTA needs from global memory is requested before the first barrier, so that the L2 round trips overlap
    const View vw = a.views[f];
    const uint32_t slot = a.frame_slot[f];
    ColIdx ci_first;
    ci_first.first = 0; ci_first.n = 0;
    if (warp < NSETS * NSPLIT) {
        const int x = g * TC + grp * NSETS + (NSPLIT > 1 ? warp % NSETS : warp);
        if (x < a.W) ci_first = a.colidx[((size_t)f * nlists + lband) * a.W + x];
    }
#ifdef DRR_PAL8
    // (passing this image as a by-value kernel parameter and copying it from the constant bank was measured: the divergent
    // LDCs made the empty-CTA time 3x longer)
    for (int i = threadIdx.x; i < 257 * 3; i += NT) s_pal[i] = a.pal_image[i];
#else
    for (int i = threadIdx.x; i < 257; i += NT) s_pal[i] = a.palette[min(i, 255)];
#endif
    if (threadIdx.x == 0) s_next = NW;
    __syncthreads();
    const int px16 = sat_i16(vw.pos_x), py16 = sat_i16(vw.pos_y); // visplanes.rs:119-120 `player.position.x as i16`

    // ---- draw: a warp takes column sets (G adjacent columns) until none is left; each lane group walks its column's span
    // list in draw order ("last writer wins, transparent texels do not write"): the kinds that always write
    // simply overwrite, the HOLES kinds skip their None texels
    // (NSPLIT > 1: a work item is a column set x one of NSPLIT row ranges of the band, so that more warps than column sets
    // have work; spans are clipped to the item's rows, which keeps the draw order where it matters: on the same pixel)
    const int prow = (b1 - b0 + NSPLIT) / NSPLIT;
    for (int item = warp; item < NSETS * NSPLIT;) {
        const int cs = NSPLIT > 1 ? item % NSETS : item, part = NSPLIT > 1 ? item / NSETS : 0;
        const int pb0 = b0 + part * prow, pb1 = min(b1, pb0 + prow - 1);
        // the G columns of a set are NSETS apart: their lane groups then store to disjoint bank ranges when they sit on the same rows
        const int c = grp * NSETS + cs, x = g * TC + c;
        ColIdx ci = ci_first; // the warp's first item was requested in the prologue
        if (item != warp) {
            ci.first = 0; ci.n = 0;
            if (x < a.W) ci = a.colidx[((size_t)f * nlists + lband) * a.W + x];
        }
        const int n = (a.dbg & 16) ? 0 : (int)(ci.n & ~COL_COVERED);
        const uint4 *__restrict__ P = reinterpret_cast<const uint4 *>(a.tparams) + (size_t)ci.first * 4;
        const uint32_t col_addr = tile_addr + 4u * (uint32_t)(c * RP);
        // uncovered pixels are (0,0,0) like the reference's zero-initialised Pixels::new (pixels.rs:10-14); a column whose
        // always-writing spans cover every row (the normal case, flagged by the bin kernel) needs no clearing
        if (!(ci.n & COL_COVERED) && !(a.dbg & 1))
            for (int r = pb0 - b0 + li; r <= pb1 - b0; r += LPG) sts_u32(col_addr + 4u * (uint32_t)r, 0u);
        uint4 ra_next = make_uint4(0u, 0u, 0u, 0u);
        if (n > 0) ra_next = P[0];
        for (int j = 0; __any_sync(0xffffffffu, j < n); ++j) {
            __syncwarp(); // a span may overwrite what another lane of the group stored for an earlier span of the column
            if (j < n) {
                const uint4 ra = ra_next;
                if (j + 1 < n) ra_next = P[4 * (j + 1)]; // the next span's first words are fetched while this one is drawn
                const int ya = max((int)(ra.x & 0xffff), pb0), yb = min((int)(ra.x >> 16), pb1);
                const uint32_t kind = ra.y & 0xffu;
                if (ya <= yb && kind != KIND_NONE) {
                    if (kind == KIND_FLAT) {
                        // a span clipped to a band stays inside the rows its flags were computed for
                        if ((ra.y & (TS_UNIT | TS_FASTDIV)) == (TS_UNIT | TS_FASTDIV))
                            tile_flat_span<LPG, true>(ra, P[4 * j + 2], ya, yb, b0, li, col_addr, a.CFY, vw.cos_a, vw.sin_a, px16, py16, flats, pal_addr, a.one, a.tex_flats);
                        else
                            tile_flat_span<LPG, false>(ra, P[4 * j + 2], ya, yb, b0, li, col_addr, a.CFY, vw.cos_a, vw.sin_a, px16, py16, flats, pal_addr, a.one, a.tex_flats);
                    } else if (kind == KIND_WALL) {
                        const uint4 rb = P[4 * j + 1], rc = P[4 * j + 2], rd = P[4 * j + 3];
                        if (ra.y & TS_POW2) tile_wall_span<LPG, false, true>(a, ra, rb, rc, rd, ya, yb, b0, li, col_addr, texels, pal_addr, a.one);
                        else tile_wall_span<LPG, false, false>(a, ra, rb, rc, rd, ya, yb, b0, li, col_addr, texels, pal_addr, a.one);
                    } else if (kind == KIND_WALL_HOLES) {
                        const uint4 rb = P[4 * j + 1], rc = P[4 * j + 2], rd = P[4 * j + 3];
                        if (ra.y & TS_POW2) tile_wall_span<LPG, true, true>(a, ra, rb, rc, rd, ya, yb, b0, li, col_addr, texels, pal_addr, a.one);
                        else tile_wall_span<LPG, true, false>(a, ra, rb, rc, rd, ya, yb, b0, li, col_addr, texels, pal_addr, a.one);
                    } else if (kind == KIND_SKY) {
                        tile_sky_span<LPG, false>(ra, ya, yb, b0, li, col_addr, a.sky_rows, texels, pal_addr);
                    } else if (kind == KIND_SKY_HOLES) {
                        tile_sky_span<LPG, true>(ra, ya, yb, b0, li, col_addr, a.sky_rows, texels, pal_addr);
                    }
                }
            }
        }
        int nx = 0;
        if (lane == 0) nx = atomicAdd(&s_next, 1);
        item = __shfl_sync(0xffffffffu, nx, 0);
    }
    __syncthreads(); // every span of the tile is in before the write-out

    // ---- write-out: Pixels::set (pixels.rs:22-30), RGB24 at 3*(y*W + x).  A row of the tile is TC*3 bytes = LPR 16-byte
    // vectors; a warp step covers RPI rows with LPR lanes each.  Vector j of a row holds bytes 16j .. 16j+15, i.e. pixels
    // (16j)/3 .. (16j+15)/3 of the tile row, starting at channel j % 3 of the first one.
    const size_t pitch = (size_t)a.W * 3;
    uint8_t *base = a.frames + (size_t)slot * a.frame_stride + (size_t)g * (TC * 3);
    const int nrows = b1 - b0 + 1;
    if (FAST_STORE) {
        constexpr int LPR = TC * 3 / 16, RPI = 32 / LPR;
        const int rl = lane / LPR, j = lane % LPR, ph = j % 3;
        uint64_t acc = 0;
        if (lane < LPR * RPI && !(a.dbg & 8)) {
            const int cb = (16 * j) / 3;
            // word m of the vector starts at byte 16j + 4m of the row: pixel q(m), channel (ph + m) % 3 -> selector; the pixel
            // pairs are (0,1) (1,2)|(2,3) (2,3)|(3,4) (4,5) relative to cb, depending on the phase
            const uint32_t s0 = wsel(ph), s1 = wsel((ph + 1) % 3), s2 = wsel((ph + 2) % 3), s3 = s0;
            const uint32_t pw = (uint32_t)(pitch >> 2);
            int r = warp * RPI + rl;
            uint32_t ta = tile_addr + 4u * (uint32_t)(cb * RP + r);
            uint32_t *wp = reinterpret_cast<uint32_t *>(base) + (size_t)(b0 + r) * pw + 4 * j;
            // checksum weight of word i is (i + 1) * C mod 2^32 (drr.h); this lane's words are i0 .. i0+3, i0 advancing by a row step
            const uint32_t C = 0x9E3779B1u;
            uint32_t k0 = ((uint32_t)(b0 + r) * pw + (uint32_t)g * (TC * 3 / 4) + 4u * (uint32_t)j + 1u) * C;
            const uint32_t kstep = (uint32_t)(NW * RPI) * pw * C;
            const size_t wstep = (size_t)(NW * RPI) * pw;
#pragma unroll 2
            for (; r < nrows; r += NW * RPI) {
                const uint32_t p0 = lds_u32(ta), p1 = lds_u32(ta + 4u * RP), p2 = lds_u32(ta + 8u * RP), p3 = lds_u32(ta + 12u * RP),
                               p4 = lds_u32(ta + 16u * RP), p5 = lds_u32(ta + 20u * RP);
                uint4 v;
                v.x = __byte_perm(p0, p1, s0);
                v.y = __byte_perm(ph == 2 ? p2 : p1, ph == 2 ? p3 : p2, s1);
                v.z = __byte_perm(ph == 0 ? p2 : p3, ph == 0 ? p3 : p4, s2);
                v.w = __byte_perm(p4, p5, s3);
                if (!(a.dbg & 2)) *reinterpret_cast<uint4 *>(wp) = v;
                if (!(a.dbg & 4)) acc += (uint64_t)v.x * k0 + (uint64_t)v.y * (k0 + C) + (uint64_t)v.z * (k0 + 2u * C) + (uint64_t)v.w * (k0 + 3u * C);
                ta += 4u * (NW * RPI);
                wp += wstep;
                k0 += kstep;
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0 && acc) atomicAdd(reinterpret_cast<unsigned long long *>(a.crc + slot), (unsigned long long)acc);
    } else {
        for (int i = threadIdx.x; i < TC * nrows; i += NT) { // generic widths: bytewise; the checksum is a separate pass
            const int c = i % TC, r = i / TC, x = g * TC + c;
            if (x >= a.W) continue;
            const uint32_t rgb = s_tile[c * RP + r];
            uint8_t *p = base + (size_t)(b0 + r) * pitch + c * 3;
            p[0] = (uint8_t)rgb;
            p[1] = (uint8_t)(rgb >> 8);
            p[2] = (uint8_t)(rgb >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
cudaError_t launch_bin(const DrawArgs &a, int frame0, int nframes, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.frame_cursor + frame0, 0, sizeof(uint32_t) * (size_t)nframes, st);
    if (e != cudaSuccess) return e;
    int maxt = 192;
    if (const char *e = getenv("DRR_BIN_THREADS")) maxt = std::max(32, std::min(BIN_THREADS, atoi(e) / 32 * 32));
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

void tile_bands(int H, int *nbands, int *band_rows) {
    // rows per band: at most 400 (a 32-column tile is then <= 51 KB: four CTAs per SM), equal bands.  Since the bin kernel writes
    // one span list per (column, band), wide-and-short tiles beat the 16-column full-height ones (tools/sweep_wide.sh).
    int max_rows = 400;
    if (const char *e = getenv("DRR_TILE_MAX_ROWS")) max_rows = std::max(32, atoi(e));
    *nbands = (H + max_rows - 1) / max_rows;
    *band_rows = (H + *nbands - 1) / *nbands;
}

void tile_config(int W, int H, int *tc, int *lpg) {
    // measured (tools/sweep_tile.sh, tools/sweep_env.sh, tools/sweep_wide.sh; profiles/r1_ab_measurements.md): 32-column tiles
    // (96-byte row segments: full 32-byte sectors) with 8 lanes per span at every resolution -- short lane groups keep the
    // warp full on short spans; 16-column tiles and other group sizes stay selectable for A/B runs
    *tc = 32;
    *lpg = 8;
    if (const char *e = getenv("DRR_TILE_COLS")) {
        const int v = atoi(e);
        if (v == 16 || v == 32) *tc = v;
    }
    if (const char *e = getenv("DRR_TILE_LPG")) {
        const int v = atoi(e);
        if (v == 8 || v == 16 || v == 32) *lpg = v;
    }
}

template <int TC, int LPG, int NT, int NSPLIT>
static cudaError_t launch_tile_t(const DrawArgs &a, int frame0, int nframes, cudaStream_t st, int *launches) {
    const int gpf = (a.W + TC - 1) / TC;
    // rows per band: the whole column while the tile stays within ~52 KB (TC 16) / ~105 KB (TC 32), else equal bands
    const int nbands = a.nbands, band_rows = a.band_rows; // tile_bands(), shared with the bin kernel
    const int want = TC == 16 ? 2 : 1;
    int RP = band_rows;
    while (RP % 32 != want) ++RP;
    const long long blocks = (long long)nframes * gpf * nbands;
    if (blocks == 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    size_t dyn = ((size_t)TC * RP * 4 + 15) / 16 * 16;
    if (const char *e = getenv("DRR_TILE_SMEM_PAD_KB")) dyn += (size_t)atoi(e) * 1024; // A/B: fewer CTAs per SM, more L1
    const bool fast = (a.W % TC) == 0;
    *launches = 1;
    cudaError_t e;
    auto prepare = [&](auto kernel) -> cudaError_t {
        static bool done = false;
        if (done) return cudaSuccess;
        done = true;
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    };
    // one CTA per tile (a persistent variant with a static tile stride was measured and lost: profiles/r1_ab_measurements.md)
    if ((long long)gpf * nbands > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    for (int f0 = 0; f0 < nframes; f0 += 65535) { // gridDim.y limit
        const dim3 grid((unsigned)(gpf * nbands), (unsigned)std::min(65535, nframes - f0));
        if (fast) {
            if ((e = prepare(drr_tile_kernel<TC, LPG, NT, NSPLIT, true>)) != cudaSuccess) return e;
            drr_tile_kernel<TC, LPG, NT, NSPLIT, true><<<grid, NT, dyn, st>>>(a, frame0 + f0, band_rows, nbands, RP);
        } else {
            if ((e = prepare(drr_tile_kernel<TC, LPG, NT, NSPLIT, false>)) != cudaSuccess) return e;
            drr_tile_kernel<TC, LPG, NT, NSPLIT, false><<<grid, NT, dyn, st>>>(a, frame0 + f0, band_rows, nbands, RP);
        }
        if (f0) ++*launches;
    }
    if (!fast) {
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch_checksum_pass(a, frame0, nframes, st, launches);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t launch_tile(const DrawArgs &a, int frame0, int nframes, cudaStream_t st, int *launches) {
    int tc, lpg;
    tile_config(a.W, a.H, &tc, &lpg);
    if (tc == 16) {
        // A/B knob DRR_TILE_SPLIT=2: 12 warps (40 registers) share a tile and each column set is cut into two row ranges, so
        // that 48 instead of 32 warps are resident per SM -- measured slower (1.30 vs 1.16 ms at 1280x800), kept for the record
        const char *e = getenv("DRR_TILE_SPLIT");
        if (e && atoi(e) == 2) {
            if (lpg == 32) return launch_tile_t<16, 32, 384, 2>(a, frame0, nframes, st, launches);
            if (lpg == 16) return launch_tile_t<16, 16, 384, 2>(a, frame0, nframes, st, launches);
            return launch_tile_t<16, 8, 384, 2>(a, frame0, nframes, st, launches);
        }
        if (lpg == 32) return launch_tile_t<16, 32, 256, 1>(a, frame0, nframes, st, launches);
        if (lpg == 16) return launch_tile_t<16, 16, 256, 1>(a, frame0, nframes, st, launches);
        return launch_tile_t<16, 8, 256, 1>(a, frame0, nframes, st, launches);
    }
    if (lpg == 32) return launch_tile_t<32, 32, 256, 1>(a, frame0, nframes, st, launches);
    if (lpg == 16) return launch_tile_t<32, 16, 256, 1>(a, frame0, nframes, st, launches);
    return launch_tile_t<32, 8, 256, 1>(a, frame0, nframes, st, launches);
}

} // namespace drr
