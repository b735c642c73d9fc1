// drr_kernels.h -- kernel argument block and launchers (shared by drr_tile.cu, drr_kernels.cu and drr_api.cu)
#pragma once
#include "drr_device.cuh"
#include <cuda.h> // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)

namespace drr {

static constexpr int TILE_THREADS = 256; // 8 warps: one per group of four adjacent screen columns of the 32-column tile
static constexpr int TILE_COLS = 32;     // 96-byte row segments of the RGB24 framebuffer: whole 32-byte sectors
static constexpr int TILE_LPG = 8;       // lanes per span: a warp draws the four columns of its group at once

// Shared memory of a tile CTA.  All of it is dynamic (the kernel declares no static shared memory), so the map starts at the
// CTA's shared window base, which drr_ctx_create probes once (DrawArgs::pal_base): the texel pool holds ABSOLUTE shared
// addresses of palette entries (pal_base + 8 * index), so a texel feeds the palette load without an add.
//   [SM_PAL, +3088)  palette image: 257 x 8 bytes (r and g as bf16 -- exact for 0..255 -- in one word, b as f32; entry 256 backs
//                    the None texel), then 257 packed 0x00BBGGRR words (sky: no lighting); staged by ONE bulk copy
//   [SM_BAR, +8)     mbarrier of that copy
//   [SM_TILE, ...)   the tile: band_rows x 128 bytes, see tile_offset()
static constexpr uint32_t PAL_ENTRY = 8;
static constexpr uint32_t SM_PAL = 0, SM_PAL_PACKED = 257 * PAL_ENTRY, SM_PAL_BYTES = 3088, SM_BAR = 3088, SM_TILE = 3200;
static constexpr uint32_t TEXEL_NONE_INDEX = 256; // palette index that stands for a None texel in the pool

// Byte offset of pixel (column c of the tile, row r of the band) in the tile.  Blocks of 8 rows (1 KB); inside a block the
// four columns of a group ("quad") are interleaved per row: 16 bytes = pixels c&3 = 0..3 of one row, so that
//   * drawing is conflict-free whatever rows the four lane groups of a warp are on: 8 lanes = 8 consecutive rows = 8
//     different 16-byte slots, the 4 groups = the 4 words of a slot;
//   * the write-out reads four pixels of a row with one LDS.128, and (rows ^ 4 for the upper four quads) a quarter warp
//     reading 4 rows x 2 halves of the tile hits eight different slots;
//   * the rows y and y + 8 of a lane are exactly 1 KB apart.
__host__ __device__ __forceinline__ uint32_t tile_offset(int c, int r) {
    const uint32_t q = (uint32_t)c >> 2;
    return (((uint32_t)r >> 3) << 10) + (q << 7) + ((((uint32_t)r & 7u) ^ (q & 4u)) << 4) + (((uint32_t)c & 3u) << 2);
}

struct DrawArgs {
    int W, H, nframes;
    int nbands, band_rows;           // rows of a tile: the whole column (nbands == 1) or equal row bands (tile_bands())
    // src/renderer/constants.rs:7-17 derived from W, H with the reference's own expressions (drr_ctx_create)
    float CFX, CFY, GCFX, ASPECT, Wf, Hf;
    int dbg;   // timing experiments only (DRR_DBG): 1 skip clearing, 2 skip framebuffer stores, 4 skip checksum, 8 skip write-out, 16 skip drawing
    float one; // always 1.0f, but opaque to ptxas: see add2_nofuse() in drr_tile.cu
    // draw lists as emitted (all frames of the batch concatenated, draw order)
    const View *views;               // [nframes]
    const uint32_t *ops;             // per frame, in call order: bit 31 = visplane, low bits = index into planes[] / segs[]
    const uint32_t *frame_op_base;   // [nframes + 1]
    const uint32_t *frame_nops;      // [nframes] ops of each frame when the lists are the front-end's per-view slabs (null: frame_op_base[f + 1] - frame_op_base[f])
    const uint32_t *frame_rec_base;  // [nframes + 1] first record slot of each frame (one slot per emitted column: an upper bound)
    const uint32_t *frame_slot;      // framebuffer slot (view index) of each recorded frame
    const SegRec *segs;
    const ColRec *cols;
    const PlaneRec *planes;
    const uint32_t *parr;            // (top, bottom) i16 pairs of the visplane columns
    // device scratch written by the bin kernel, read by the tile kernel
    uint32_t *frame_cursor;          // [nframes] records handed out so far (zeroed before the bin kernel)
    ColIdx *colidx;                  // [nframes * nlists * W], nlists = nbands when nbands <= MAX_LIST_BANDS, else 1
    void *tparams;                   // one 64-byte decoded record per (op, column) that survives clipping, see drr_tile.cu
    // assets
    const uint16_t *texels;          // bitmap pool: column-major, pow2 column pitch; values = pal_base + 8 * palette index (256 = None)
    const uint8_t *flats;            // 4096 bytes per flat slot (palette indices)
    const BitmapRec *bitmaps;
    const uint32_t *pal_image;       // the palette exactly as the tile kernel's shared memory holds it (SM_PAL_BYTES, 16-byte aligned)
    uint32_t pal_base;               // shared window address of a CTA's dynamic shared memory (probed at drr_ctx_create)
    const uint8_t *sky_rows;         // sky texture row of every screen row
    uint32_t sky_base;               // texel index of the 256x128 sky bitmap
    uint8_t *frames;                 // framebuffers, frame_stride bytes apart, RGB24 row-major
    uint64_t frame_stride;
    uint64_t *crc;                   // per-slot checksum accumulators (zeroed before the draw)
};

// Every launcher works on the frame range [frame0, frame0 + nframes) of the uploaded batch.
cudaError_t launch_bin(const DrawArgs &a, int frame0, int nframes, cudaStream_t st);
// fbmap: the framebuffers as a 3-D u8 tensor (x bytes, row, slot) with a 96-byte x 8-row box, or null when W % 32 != 0
cudaError_t launch_tile(const DrawArgs &a, const CUtensorMap *fbmap, int frame0, int nframes, cudaStream_t st, int *launches);
cudaError_t probe_shared_base(uint32_t *base); // shared window address of a CTA's dynamic shared memory on the current device
cudaError_t launch_sky_rows(uint8_t *rows, int H, cudaStream_t st);
// CRC-32 (zlib polynomial) of nframes resident framebuffers starting at slot `first`: out[k] = crc32 of frames + (first + k) * stride
cudaError_t launch_crc32(const uint8_t *frames, uint64_t frame_stride, uint64_t nbytes, int first, int nframes, uint32_t *chunk_scratch,
                         uint32_t *out, cudaStream_t st);
size_t crc32_scratch_words(uint64_t nbytes, int nframes);
cudaError_t launch_checksum_pass(const DrawArgs &a, int frame0, int nframes, cudaStream_t st, int *launches);
void tile_config(int W, int H, int *tc, int *lpg);
void tile_bands(int H, int max_rows, int *nbands, int *band_rows); // how the tile kernel cuts a column into row bands (equal bands of at most 400 rows, a multiple of 8 when H is)
static constexpr int MAX_LIST_BANDS = 8;             // up to this many bands the bin kernel writes one span list per (column, band)
// ---- device front-end (drr_frontend.cu / drr_frontend.cuh) ----------------------------------------------------------
namespace fe {
struct Map;
struct ViewIn;
struct Counts;
struct Bases;
struct Out;
struct Caps;
struct Slabs;
} // namespace fe
// Shape of the front-end kernel (one warp per viewpoint): warps per CTA and resident CTAs per SM the register budget is set for.
// The kernel wants ~146 registers.  Measured per 4096 viewpoints of walk320, two warps per CTA: 16 CTAs per SM (64 registers,
// spills) 0.834 ms, 12 (80) 0.736, 10 (96) 0.697, 8 (128) 0.644, 6 (146) 0.720 -- fewer, fatter warps win as long as the per-view
// arrays sit in shared memory; four warps per CTA at the same 128 registers (neighbouring viewpoints walk the same part of the
// map: they share L1) 0.618 ms, things640 1.90 -> 1.79 ms.  With the arrays in global scratch (wide screens: the stress map at
// 1920x1200) more warps hide more than the spills cost: 64 registers 56.1 ms, 128 registers 60.5 ms per 8192 viewpoints.
#ifndef DRR_FE_WARPS
#define DRR_FE_WARPS 4
#endif
#ifndef DRR_FE_MIN_BLOCKS
#define DRR_FE_MIN_BLOCKS 4
#endif
#ifndef DRR_FE_WARPS_GLOBAL
#define DRR_FE_WARPS_GLOBAL 2
#endif
#ifndef DRR_FE_MIN_BLOCKS_GLOBAL
#define DRR_FE_MIN_BLOCKS_GLOBAL 16
#endif
static constexpr int FE_WARPS = DRR_FE_WARPS, FE_MIN_BLOCKS = DRR_FE_MIN_BLOCKS;                             // per-view arrays in shared memory
static constexpr int FE_WARPS_GLOBAL = DRR_FE_WARPS_GLOBAL, FE_MIN_BLOCKS_GLOBAL = DRR_FE_MIN_BLOCKS_GLOBAL; // ... in global scratch
static constexpr int FE_WARPS_THINGS = 8, FE_MIN_BLOCKS_THINGS = 2;                                           // ... in shared memory, map objects drawn
struct FeScratch {                    // per-viewpoint working state of the front-end, W entries per viewpoint each
    uint8_t *hor_ocl;
    int16_t *floor_ocl, *ceil_ocl;
    uint32_t *rows;                   // 2 * W per viewpoint: the (top, bottom) rows of the two visplanes being accumulated
    int32_t *order;                   // nsegs per viewpoint: the segs in the view's BSP order (then the node side bits)
    void *pre;                        // nsegs records (fe::SegPre, 32 bytes) per viewpoint, written by drr_fe_pre_kernel; null = evaluate in the walk
    uint8_t *pre_code;                // nsegs codes per viewpoint
    // masked phase only (null otherwise); per viewpoint: cap_* entries each, cap_dsegs part indices, cap_mos draw-order entries
    void *renders, *allcols, *dsegs, *mos;
    int32_t *mo_order;
    int32_t *dseg_part;
    uint32_t cap_renders, cap_allcols, cap_dsegs, cap_mos;
};
// emit == false: count pass (writes counts[0..n)); emit == true: writes the lists at the offsets in bases[0..n), or -- when
// slab.ops != 0, single-pass mode -- into per-view slabs of those capacities, leaving counts[0..n) for launch_fe_compact
// the stateless part of process_seg for every (viewpoint, seg) pair: fills s.pre / s.pre_code
cudaError_t launch_fe_pre(const fe::Map &m, const fe::ViewIn *views, int n, const FeScratch &s, cudaStream_t st);
cudaError_t launch_frontend(bool emit, const fe::Map &m, const fe::ViewIn *views, const fe::Bases *bases, fe::Counts *counts, int n,
                            const FeScratch &s, const fe::Out &out, const fe::Caps &slab, cudaStream_t st);
cudaError_t launch_fe_compact(const fe::Slabs &sl, const fe::Counts *counts, const fe::Bases *bases, int n, const fe::Out &dst, cudaStream_t st);
// frame f's View out of the slab array (indexed by viewpoint = frame_slot[f] - first_view_idx) into the dense one the draw kernels index by frame
cudaError_t launch_fe_gather_views(const View *slab_views, const uint32_t *frame_slot, int first_view_idx, int nframes, View *dst, cudaStream_t st);

cudaError_t launch_fastdiv_check(int mode, long long n0, long long n1, float CFY, int H, uint32_t lo, uint32_t stride,
                                 unsigned long long *d_bad, float *d_first, cudaStream_t st);

} // namespace drr
