// drr_device.cuh -- device data layout and exact-arithmetic helpers shared by the kernels.
//
// Arithmetic contract: every f32 operation of the reference is reproduced with an explicitly
// rounded intrinsic (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn), which nvcc never contracts into an FMA and
// never reassociates, regardless of -fmad / -use_fast_math.  Float->int casts follow Rust `as` (truncate, saturate,
// NaN -> 0); i16 arithmetic wraps (release build).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace drr {

// ---- span kinds -------------------------------------------------------------------------------------------------
enum : uint32_t {
    KIND_WALL = 0,       // wall / sprite column whose bitmap has no transparent texel (always writes)
    KIND_WALL_HOLES = 1, // bitmap with None texels (masked mid-textures, sprites, holey walls): may skip pixels
    KIND_FLAT = 2,       // floor / ceiling
    KIND_SKY = 3,        // sky, opaque sky bitmap
    KIND_SKY_HOLES = 4,  // sky bitmap with None texels
};

// One emitted column of one op after clipping to the screen, in draw order (what the bin kernel turns into a TileSpan).
// On the device it only ever exists in registers; this struct is the host-side mirror used by the CPU tests
// (drr_test_list) to check the device binning. 16 bytes.
struct Span {
    uint16_t y0, y1; // inclusive screen rows
    uint16_t x;      // screen column
    uint8_t kind;
    uint8_t pad;
    uint32_t op;     // index into segs[] (wall kinds) or planes[] (flat / sky kinds), global over the batch
    int16_t top_y, bottom_y; // BitmapColumn.top_y / bottom_y (wall kinds only)
};
static_assert(sizeof(Span) == 16, "Span layout");

// Device copy of drr_seg_hdr (bitmap id resolved to a slot) plus where its column records are. 80 bytes.
struct SegRec {
    uint32_t bitmap_slot;
    int16_t light_level;
    int16_t phase;
    float lsx, lsy, lex, ley;
    float start_offset;
    int32_t start_x, end_x;
    float bottom_height, top_height;
    int16_t offset_x, offset_y;
    uint32_t cols_first; // index of the first ColRec; the records' x is strictly increasing (drr_emit_columns splits otherwise)
    uint32_t n;          // number of ColRec
    int16_t x0, x1;      // x of the first / last record
    // the bitmap's record (BitmapRec of bitmap_slot), copied in at emit time so that the bin kernel needs no dependent load
    uint32_t tex_base;
    int16_t tex_w, tex_h;
    uint32_t tex_opaque;
    uint32_t pad[2];
};
static_assert(sizeof(SegRec) == 80, "SegRec layout");

// == drr_col (BitmapColumn, bitmap_render.rs:19-25). 10 bytes, 2-byte aligned.
struct ColRec {
    int16_t x, clipped_top_y, clipped_bottom_y, bottom_y, top_y;
};
static_assert(sizeof(ColRec) == 10, "ColRec layout");

struct PlaneRec { // 16 bytes
    int16_t flat_slot; // -1 = sky
    int16_t height;
    int16_t light_level;
    int16_t left, right;
    int16_t kind;       // KIND_FLAT / KIND_SKY / KIND_SKY_HOLES
    uint32_t arr_first; // index into the (top, bottom) i16 pair pool of the entry for x == left
};
static_assert(sizeof(PlaneRec) == 16, "PlaneRec layout");

struct ColIdx { // per (frame, screen column). 8 bytes
    uint32_t first; // first TileSpan of the column (global index); the column's spans follow in draw order
    uint32_t n;
};
static_assert(sizeof(ColIdx) == 8, "ColIdx layout");

struct BitmapRec { // 12 bytes
    uint32_t base; // index into the u16 texel pool (column-major, pow2 column pitch, palette byte offsets, 4096 = None)
    int16_t w, h;
    uint32_t opaque;
};

struct View { // == drr_view, 24 bytes
    float pos_x, pos_y, floor_height, angle, cos_a, sin_a;
};

// ---- Rust scalar semantics ----------------------------------------------------------------------------------------
// `f as i16`: cvt.rzi saturates to the destination range and maps NaN to 0 (PTX ISA, cvt: "float-to-integer
// conversions ... clamped to the destination range; NaN -> 0").
__device__ __forceinline__ int sat_i16(float f) {
    int r; // 32-bit destination: the s16 result arrives sign-extended (F2I.S16), no widening instruction needed
    asm("cvt.rzi.s16.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ uint32_t sat_u8(float f) { // `f as u8`
    uint32_t r = __float2uint_rz(f);                  // saturates below at 0, NaN -> 0
    return r > 255u ? 255u : r;
}
__device__ __forceinline__ int wrap16(int v) { return (int)(short)v; } // i16 wrapping result of an i32 computation

// The reference's idiom for a non-negative remainder (bitmap_render.rs:245-248, 260-263, visplanes.rs:56-58):
//     if t < 0 { t += n * (1 - t / n) }  t %= n;       all in wrapping i16
__device__ __forceinline__ int rust_wrap_mod16(int t, int n) {
    if (t < 0) {
        int q = wrap16(1 - t / n);
        t = wrap16(t + wrap16(n * q));
    }
    return t % n; // may be negative after an i16 overflow: the reference would index out of bounds and panic
}

// diminish_color's factor (bitmap_render.rs:190-201): light/255 - dist * (1/4096), clamped below at 0.
__device__ __forceinline__ float light_factor(float light_over_255, int dist_i16) {
    float f = __fsub_rn(light_over_255, __fmul_rn((float)dist_i16, 0.000244140625f));
    return f < 0.0f ? 0.0f : f;
}
// (c as f32 * factor) as u8 per channel (bitmap_render.rs:203-207), packed 0x00BBGGRR. pal = (r, g, b) as floats.
__device__ __forceinline__ uint32_t lit_rgb(float4 pal, float factor) {
    uint32_t r = sat_u8(__fmul_rn(pal.x, factor));
    uint32_t g = sat_u8(__fmul_rn(pal.y, factor));
    uint32_t b = sat_u8(__fmul_rn(pal.z, factor));
    return r | (g << 8) | (b << 16);
}

// Position-weighted linear checksum (drr.h: drr_read_checksums).  The frame's little-endian u32 words (zero-padded) are taken
// in groups of 12 (48 bytes = 16 pixels): s_g = sum_j w[12g + j] * ((2j + 1) * C) mod 2^32, checksum = sum_g s_g * ((g + 1) * C
// mod 2^32) mod 2^64, C = 0x9E3779B1.  Every multiplier is odd, so any change of a single word changes the sum.  (A lane of the
// tile kernel's write-out holds exactly one group: twelve 32-bit multiply-adds and one wide one, where a weight per word cost
// twelve wide ones.)
static constexpr uint32_t CK_C = 0x9E3779B1u;
static constexpr int CK_GROUP = 12;
__host__ __device__ __forceinline__ uint32_t checksum_word_weight(int j) { return (uint32_t)(2 * j + 1) * CK_C; }
__host__ __device__ __forceinline__ uint64_t checksum_group_term(uint32_t s, uint64_t group) {
    const uint32_t k = (uint32_t)(group + 1u) * CK_C; // never 0 for group + 1 < 2^32 (the multiplier is odd)
    return (uint64_t)s * (uint64_t)k;
}

} // namespace drr
