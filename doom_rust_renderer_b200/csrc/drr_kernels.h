// drr_kernels.h -- kernel argument block and launchers (shared by drr_kernels.cu and drr_api.cu)
#pragma once
#include "drr_device.cuh"

namespace drr {

static constexpr int MARCH_THREADS = 128;
#ifndef DRR_MARCH_MIN_BLOCKS
#define DRR_MARCH_MIN_BLOCKS 6
#endif
static constexpr int MARCH_MIN_BLOCKS = DRR_MARCH_MIN_BLOCKS; // occupancy target: 6 CTAs x 4 warps per SM (<= 80 registers)

static constexpr int TILE_THREADS = 256;
#ifndef DRR_TILE_MIN_BLOCKS
#define DRR_TILE_MIN_BLOCKS 4
#endif
static constexpr int TILE_MIN_BLOCKS = DRR_TILE_MIN_BLOCKS;

struct DrawArgs {
    int W, H, nframes;
    int colmajor; // texel pool layout: 1 = column-major (tile kernel), 0 = row-major (march kernel)
    // src/renderer/constants.rs:7-17 derived from W, H with the reference's own expressions (see make_constants())
    float CFX, CFY, GCFX, ASPECT, Wf, Hf;
    float one; // always 1.0f, but opaque to ptxas: see add2_nofuse() in drr_kernels.cu
    const View *views;
    const SegRec *segs;
    const PlaneRec *planes;
    const Span *spans;
    const uint32_t *frame_span_base; // nframes + 1 entries
    const uint32_t *frame_slot;      // framebuffer slot (view index) of each recorded frame
    const ColIdx *colidx;            // nframes * W entries
    SpanParams *params;              // one per span (march kernel)
    void *tparams;                   // one 64-byte decoded record per span (tile kernel, drr_tile.cu)
    const uint8_t *sky_rows;         // sky texture row of every screen row (tile kernel)
    const uint16_t *texels;          // bitmap pool: march = row-major, pow2 row pitch, palette index, 0x8000 = None;
                                     //              tile  = column-major, pow2 column pitch, palette byte offset (index*16), 4096 = None
    const uint8_t *flats;            // 4096 bytes per flat slot
    const BitmapRec *bitmaps;
    const float4 *palette;           // 256 x (r, g, b as f32, packed 0x00BBGGRR bits)
    uint32_t sky_base;               // texel index of the 256x128 sky bitmap
    uint8_t *frames;                 // framebuffers, frame_stride bytes apart, RGB24 row-major
    uint64_t frame_stride;
    uint64_t *crc;                   // per-frame checksum accumulators (zeroed before the launch)
};

cudaError_t launch_span_setup(const DrawArgs &a, uint32_t nspans, cudaStream_t st);
cudaError_t launch_march(const DrawArgs &a, cudaStream_t st, int *launches);
cudaError_t launch_tile(const DrawArgs &a, cudaStream_t st, int *launches);
cudaError_t launch_tile_setup(const DrawArgs &a, uint32_t nspans, cudaStream_t st);
cudaError_t launch_sky_rows(uint8_t *rows, int H, cudaStream_t st);
cudaError_t launch_checksum_pass(const DrawArgs &a, cudaStream_t st, int *launches);
void tile_config(int W, int H, int *tc, int *lpg);
cudaError_t launch_fastdiv_check(int mode, long long n0, long long n1, float CFY, int H, uint32_t lo, uint32_t stride,
                                 unsigned long long *d_bad, float *d_first, cudaStream_t st);

} // namespace drr
