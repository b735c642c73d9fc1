// drr_api.cu -- the C ABI of include/drr.h: context, asset upload, draw-list recording, upload and launches.
//
// Data layout in HBM (all frames of a batch concatenated, see drr_device.cuh):
//   as emitted:  views[frame] 24 B | ops[] 4 B (call order) | segs[] 80 B | cols[] 10 B | planes[] 16 B | (top, bottom) pairs 4 B
//   device scratch: colidx[frame][x] 8 B and one decoded 64-byte record per (op, column), both written by the bin kernel
//   framebuffers: max_views x (W*H*3 B, RGB24 row-major == Pixels.pixels, src/renderer/pixels.rs:5-14)
//   assets: u16 texel pool (column-major, pow2 column pitch; values = shared address of the palette entry in a tile CTA,
//           entry 256 = None), u8 flat pool (4096 B per flat), the palette image a tile CTA stages
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/drr.h"
#include "drr_kernels.h"
#include "drr_frontend.cuh"
#include <cmath>
#include <chrono>
#include <type_traits>

using namespace drr;

namespace {

thread_local std::string g_create_error;

template <class T>
struct PinnedVec { // growable pinned staging buffer (cudaHostAlloc), so H2D copies are truly asynchronous
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = true; // false only in the CPU-test recording context (drr_test_ctx_create_host_only)
    ~PinnedVec() { release(p); }
    void release(T *q) {
        if (!q) return;
        if (pinned) cudaFreeHost(q); else free(q);
    }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t nc = std::max<size_t>(want, cap ? cap * 2 : 1024);
        T *q = nullptr;
        if (pinned) {
            if (cudaHostAlloc((void **)&q, nc * sizeof(T), cudaHostAllocDefault) != cudaSuccess) return false;
        } else if (!(q = (T *)malloc(nc * sizeof(T)))) {
            return false;
        }
        if (n) memcpy(q, p, n * sizeof(T));
        release(p);
        p = q;
        cap = nc;
        return true;
    }
    bool push(const T &v) {
        if (n == cap && !reserve(n + 1)) return false;
        p[n++] = v;
        return true;
    }
    void clear() { n = 0; }
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t want) {
        if (want <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t nc = want + want / 4 + 16;
        cudaError_t e = cudaMalloc((void **)&p, nc * sizeof(T));
        if (e == cudaSuccess) cap = nc;
        return e;
    }
};

} // namespace

// Draw lists exactly as the host emitted them, all frames concatenated.  A context records into its own (pinned staging,
// so that the H2D copies are asynchronous); a drr_recorder records into a private one (plain memory) that is later appended
// to a context -- that is how worker threads run the front-end in parallel.
struct Lists {
    PinnedVec<View> views;
    PinnedVec<uint32_t> ops;            // per frame, call order: bit 31 = visplane, low bits = index into planes / segs
    PinnedVec<uint32_t> frame_op_base;  // frames + 1
    PinnedVec<uint32_t> frame_rec_base; // frames + 1: records (columns that survive clipping) before each frame
    PinnedVec<uint32_t> frame_slot;
    PinnedVec<SegRec> segs;
    PinnedVec<ColRec> cols;
    PinnedVec<PlaneRec> planes;
    PinnedVec<uint32_t> parr;           // (top, bottom) i16 pairs
    std::vector<uint32_t> frame_seg_base, frame_col_base, frame_plane_base, frame_parr_base; // frames + 1 each (chunked upload)
    uint64_t rec_count = 0; // columns that survive clipping, all frames recorded so far (statistics)
    uint64_t rec_cap = 0;   // screen columns inside the x ranges of all ops recorded so far: what the bin kernel may reserve (>= rec_count)
    // frame being recorded
    bool in_frame = false;
    size_t ops_n0 = 0, segs_n0 = 0, cols_n0 = 0, planes_n0 = 0, parr_n0 = 0; // list sizes at drr_frame_begin (for drr_frame_abort)
    uint64_t rec0 = 0, cap0 = 0;
    drr_stats stats0{};
    int cur_slot = -1;

    drr_stats stats{};
    void set_pinned(bool on) {
        views.pinned = ops.pinned = frame_op_base.pinned = frame_rec_base.pinned = frame_slot.pinned = on;
        segs.pinned = cols.pinned = planes.pinned = parr.pinned = on;
    }
    void clear_lists() {
        views.clear(); ops.clear(); frame_op_base.clear(); frame_rec_base.clear(); frame_slot.clear();
        segs.clear(); cols.clear(); planes.clear(); parr.clear();
        frame_seg_base.clear(); frame_col_base.clear(); frame_plane_base.clear(); frame_parr_base.clear();
        rec_count = rec_cap = 0;
        in_frame = false;
    }
};

// Device front-end (drr_fe_*): the flattened map (host mirror with ids resolved to slots + device copy), per-view scratch
// and the count / offset tables of the last batch.
struct FeState {
    bool have_map = false;
    uint32_t map_id = 0; // changes with every drr_fe_upload_map (drr_fe_map_id)
    std::vector<fe::Thing> things;
    DevBuf<fe::Thing> d_things;
    DevBuf<fe::RenderRec> d_renders;
    DevBuf<uint16_t> d_allcols;
    DevBuf<SegRec> d_dsegs;
    DevBuf<fe::MoRec> d_mos;
    DevBuf<int32_t> d_mo_order;
    DevBuf<int32_t> d_dseg_part;
    std::vector<fe::Node> nodes;
    std::vector<fe::SubSector> ssectors;
    std::vector<fe::NodeUp> node_up; // empty when the node lump is not a proper tree
    std::vector<int32_t> ss_up;
    int nord = 0;
    DevBuf<fe::NodeUp> d_node_up;
    DevBuf<int32_t> d_ss_up;
    std::vector<fe::Seg> segs;
    std::vector<fe::Line> lines;
    std::vector<fe::Side> sides;
    std::vector<fe::Sector> sectors;
    std::vector<fe::Bitmap> bitmaps; // indexed by the caller's bitmap id
    DevBuf<fe::Node> d_nodes;
    DevBuf<fe::SubSector> d_ssectors;
    DevBuf<fe::Seg> d_segs;
    DevBuf<fe::Line> d_lines;
    DevBuf<fe::Side> d_sides;
    DevBuf<fe::Sector> d_sectors;
    DevBuf<fe::Bitmap> d_bitmaps;
    DevBuf<fe::ViewIn> d_views_in;
    DevBuf<fe::Counts> d_counts;
    DevBuf<fe::Bases> d_bases;
    DevBuf<uint8_t> d_hor;
    DevBuf<int16_t> d_focl, d_cocl;
    DevBuf<uint32_t> d_rows;
    DevBuf<int32_t> d_order;
    DevBuf<fe::SegPre> d_pre; // the stateless kernel's records and codes, nsegs per viewpoint
    DevBuf<uint8_t> d_pre_code;
    PinnedVec<fe::ViewIn> h_views_in;
    PinnedVec<fe::Counts> h_counts;
    PinnedVec<fe::Bases> h_bases;
    DevBuf<View> d_sl_views; // single-pass mode: the per-view slabs
    DevBuf<uint32_t> d_sl_ops, d_sl_parr;
    DevBuf<SegRec> d_sl_segs;
    DevBuf<uint16_t> d_sl_cols;
    DevBuf<PlaneRec> d_sl_planes;
    bool single_pass = false; // how the last batch ran
    // slab draw: the last batch's lists are still the per-view slabs (the draw kernels index them through d_slab_meta); dense lists
    // exist only after compact_now() (the test accessor that downloads them)
    bool slab_draw = false, dense_valid = false;
    fe::Slabs sl_last{};
    int n_last = 0, first_last = 0;
    PinnedVec<uint32_t> h_slab_meta; // [nf + 1] first op of each frame in the slab array, then [nf] its op count
    DevBuf<uint32_t> d_slab_meta;
    uint32_t scratch_boost = 1; // x4 whenever a batch outgrew the masked phase's working arrays
    uint32_t slab_boost = 1;  // doubled (up to 8) whenever a batch outgrew its slabs: the next batch of the context gets more room
    float count_ms = 0.0f, emit_ms = 0.0f;
    uint64_t device_list_bytes = 0; // size of the lists the last drr_fe_emit_views wrote on the device
};

// Tuning / diagnostic knobs: each has an environment variable that is read ONCE, when the context is created (never on
// the launch path); drr_set_knob() changes one afterwards.  0 / false = the built-in default.
struct Knobs {
    int dbg = 0;                 // DRR_DBG: timing experiments of a -DDRR_DBG_KNOBS build only (drr_tile.cu)
    int submit_chunks = 0;       // DRR_SUBMIT_CHUNKS: upload chunks of drr_submit (default: ~4 MB of lists each, at most 8)
    int submit_one_stream = 0;   // DRR_SUBMIT_ONE_STREAM: all chunks on the context's stream
    int submit_trace = 0;        // DRR_SUBMIT_TRACE: per-chunk time line on stderr
    int tile_max_rows = 0;       // DRR_TILE_MAX_ROWS: rows per tile band (default 400; A/B runs)
    int fe_trace = 0;            // DRR_FE_TRACE: host-side time line of drr_fe_emit_views on stderr
    int fe_two_pass = 0;         // DRR_FE_TWO_PASS: count pass + emit pass instead of slabs + compaction
    int fe_compact = 0;          // DRR_FE_COMPACT: copy the slabs into dense lists before drawing (default: the draw kernels read the slabs)
    int fe_slab_div = 0;         // DRR_FE_SLAB_DIV: shrink the per-view slabs (tests: provoke the fallback)
    int fe_cap_renders = 0, fe_cap_dsegs = 0, fe_cap_allcols_per_w = 0; // DRR_FE_CAP_*: masked phase working arrays
    struct Name { const char *env, *name; int Knobs::*field; };
    static const Name *table(size_t *n) {
        static const Name t[] = {{"DRR_DBG", "dbg", &Knobs::dbg}, {"DRR_SUBMIT_CHUNKS", "submit_chunks", &Knobs::submit_chunks},
                                 {"DRR_SUBMIT_ONE_STREAM", "submit_one_stream", &Knobs::submit_one_stream}, {"DRR_SUBMIT_TRACE", "submit_trace", &Knobs::submit_trace},
                                 {"DRR_TILE_MAX_ROWS", "tile_max_rows", &Knobs::tile_max_rows}, {"DRR_FE_TRACE", "fe_trace", &Knobs::fe_trace}, {"DRR_FE_TWO_PASS", "fe_two_pass", &Knobs::fe_two_pass}, {"DRR_FE_COMPACT", "fe_compact", &Knobs::fe_compact},
                                 {"DRR_FE_SLAB_DIV", "fe_slab_div", &Knobs::fe_slab_div}, {"DRR_FE_CAP_RENDERS", "fe_cap_renders", &Knobs::fe_cap_renders},
                                 {"DRR_FE_CAP_DSEGS", "fe_cap_dsegs", &Knobs::fe_cap_dsegs}, {"DRR_FE_CAP_ALLCOLS_PER_W", "fe_cap_allcols_per_w", &Knobs::fe_cap_allcols_per_w}};
        *n = sizeof(t) / sizeof(t[0]);
        return t;
    }
    void read_environment() {
        size_t n;
        const Name *t = table(&n);
        for (size_t i = 0; i < n; i++)
            if (const char *e = getenv(t[i].env)) this->*(t[i].field) = std::max(1, atoi(e)); // (set but "0" or empty counts as 1, as before)
    }
};

struct drr_ctx : Lists {
    FeState fes;
    Knobs knobs;
    bool device_lists = false; // the current batch's lists were written on the device (drr_fe_emit_views): nothing to upload
    int W = 0, H = 0, device = 0, max_views = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    float CFX, CFY, GCFX, ASPECT;

    // assets (host mirror + device)
    float4 pal[256];
    bool have_pal = false, assets_dirty = true;
    std::unordered_map<int, int> bitmap_slot, flat_slot;
    std::vector<BitmapRec> bitmaps;
    std::vector<uint16_t> texel_pool;
    std::vector<uint8_t> flat_pool;
    int sky_slot = -1;
    DevBuf<uint16_t> d_texels;
    DevBuf<uint8_t> d_flats;
    DevBuf<BitmapRec> d_bitmaps;
    uint32_t pal_image[SM_PAL_BYTES / 4] = {}; // the palette exactly as the tile kernel's shared memory holds it
    DevBuf<uint32_t> d_pal_image;
    uint32_t pal_base = 0x400; // shared window address of a tile CTA's dynamic shared memory (probed by drr_ctx_create)
    CUtensorMap fbmap = {};    // the framebuffers as a (x bytes, row, slot) u8 tensor with a 96-byte x 8-row box (TMA write-out)
    bool have_fbmap = false;

    // device copies of the recorded lists (the lists themselves: struct Lists above)
    DevBuf<View> d_views;
    DevBuf<uint32_t> d_ops, d_frame_op_base, d_frame_rec_base, d_frame_slot, d_parr, d_frame_cursor;
    DevBuf<SegRec> d_segs;
    DevBuf<uint16_t> d_cols; // ColRec is 10 bytes, 2-byte aligned
    DevBuf<PlaneRec> d_planes;
    DevBuf<ColIdx> d_colidx;
    DevBuf<uint4> d_tparams; // 4 x uint4 per record
    uint8_t *d_sky_rows = nullptr;
    size_t uploaded_frames = 0;
    std::vector<int> slot_to_frame; // view slot -> recorded frame (or -1)

    uint8_t *d_frames = nullptr;
    uint64_t frame_stride = 0;
    uint64_t *d_crc = nullptr;
    PinnedVec<uint64_t> h_crc;

    // host-side reference binning (test infrastructure: drr_test_list which = 3, 4), computed on demand
    std::vector<Span> t_spans;
    std::vector<ColIdx> t_colidx;

    cudaStream_t cstream = nullptr;      // copy stream of the pipelined drr_submit
    cudaStream_t stream2 = nullptr;      // second compute stream: odd chunks run here so that a chunk's bin kernel overlaps the previous chunk's draw tail
    cudaEvent_t ev_stream2 = nullptr;
    std::vector<cudaEvent_t> chunk_ev;   // "chunk uploaded" events
    cudaEvent_t ev_lists_free = nullptr; // recorded on `stream` after the last kernel that reads the device lists
    cudaEvent_t ev_h2d = nullptr;        // recorded after the last asynchronous copy that reads the pinned host lists
    bool h2d_pending = false;            // such copies may still be running: the host lists must not be touched (see lists_writable)

    std::vector<cudaEvent_t> prof_ev; // 3 events per profiled drr_draw: before setup, between, after march
    int prof_steps = 0;
    bool host_only = false; // CPU-test recording context: records and bins, can never draw
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

#define CTX_CHECK(ctx) \
    if (!(ctx)) return DRR_E_INVALID
// entry points that touch CUDA: the context's device becomes the calling thread's current device first (another context,
// or the caller itself, may have switched it since the last call)
#define CTX_DEV(ctx)                                                                        \
    CTX_CHECK(ctx);                                                                         \
    if (!(ctx)->host_only) {                                                                \
        cudaError_t e__ = cudaSetDevice((ctx)->device);                                     \
        if (e__ != cudaSuccess) {                                                           \
            (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e__);          \
            return DRR_E_CUDA;                                                              \
        }                                                                                   \
    }
#define CU(ctx, call)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            return e__ == cudaErrorMemoryAllocation ? DRR_E_NOMEM : DRR_E_CUDA;             \
        }                                                                                   \
    } while (0)

static int fail(drr_ctx *ctx, int code, const std::string &msg) {
    ctx->err = msg;
    return code;
}

// drr_upload_lists / drr_submit queue asynchronous copies out of the context's pinned host lists and return; before those
// lists are cleared, grown (reallocated) or appended to, the copies must have finished reading them.
static int lists_writable(drr_ctx *ctx) {
    if (!ctx->h2d_pending) return DRR_OK;
    CU(ctx, cudaEventSynchronize(ctx->ev_h2d));
    ctx->h2d_pending = false;
    return DRR_OK;
}

extern "C" {

const char *drr_error_name(int code) {
    switch (code) {
    case DRR_OK: return "DRR_OK";
    case DRR_E_INVALID: return "DRR_E_INVALID";
    case DRR_E_STATE: return "DRR_E_STATE";
    case DRR_E_CUDA: return "DRR_E_CUDA";
    case DRR_E_NOMEM: return "DRR_E_NOMEM";
    case DRR_E_ASSET: return "DRR_E_ASSET";
    case DRR_E_IO: return "DRR_E_IO";
    case DRR_E_PANIC: return "DRR_E_PANIC";
    default: return "DRR_E_UNKNOWN";
    }
}

const char *drr_last_error(const drr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int drr_ctx_create(int width, int height, int device_ordinal, int max_views, drr_ctx **out) {
    if (!out) return DRR_E_INVALID;
    *out = nullptr;
    if (width <= 0 || height <= 0 || width > 16384 || height > 16384 || max_views <= 0) {
        g_create_error = "drr_ctx_create: bad width/height/max_views";
        return DRR_E_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { // no CPU fallback, by design
        g_create_error = std::string("drr_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); libdrr has no CPU path";
        return DRR_E_CUDA;
    }
    if (device_ordinal < 0 || device_ordinal >= ndev) {
        g_create_error = "drr_ctx_create: bad device ordinal";
        return DRR_E_INVALID;
    }
    e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return DRR_E_CUDA;
    }
    drr_ctx *c = new drr_ctx();
    c->knobs.read_environment();
    c->W = width;
    c->H = height;
    c->device = device_ordinal;
    c->max_views = max_views;
    // src/renderer/constants.rs:7-17, same expressions, same evaluation order
    c->ASPECT = 200.0f / 240.0f;
    const float gsw = (float)(uint32_t)width / c->ASPECT;
    c->GCFX = gsw / 2.0f;
    c->CFX = (float)(uint32_t)width / 2.0f;
    c->CFY = (float)(uint32_t)height / 2.0f;
    c->slot_to_frame.assign(max_views, -1);
    c->frame_stride = ((uint64_t)width * height * 3 + 255) / 256 * 256;
    auto bail = [&](const char *what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        drr_ctx_destroy(c);
        return ce == cudaErrorMemoryAllocation ? DRR_E_NOMEM : DRR_E_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    c->own_stream = true;
    if ((e = cudaMalloc((void **)&c->d_frames, c->frame_stride * (uint64_t)max_views)) != cudaSuccess) return bail("cudaMalloc(framebuffers)", e);
    if ((e = cudaMalloc((void **)&c->d_crc, sizeof(uint64_t) * (size_t)max_views)) != cudaSuccess) return bail("cudaMalloc(crc)", e);
    if ((e = probe_shared_base(&c->pal_base)) != cudaSuccess) return bail("shared memory probe", e);
    if (c->pal_base % 128 != 0 || c->pal_base + SM_PAL + (TEXEL_NONE_INDEX + 1) * PAL_ENTRY > 0xffffu) {
        g_create_error = "drr_ctx_create: unexpected shared window base " + std::to_string(c->pal_base);
        drr_ctx_destroy(c);
        return DRR_E_CUDA;
    }
    if (width % TILE_COLS == 0 && height % 8 == 0) { // the TMA write-out's view of the framebuffers
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if ((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres)) != cudaSuccess || !fn) return bail("cuTensorMapEncodeTiled lookup", e != cudaSuccess ? e : cudaErrorUnknown);
        const cuuint64_t dims[3] = {(cuuint64_t)width * 3, (cuuint64_t)height, (cuuint64_t)max_views};
        const cuuint64_t strides[2] = {(cuuint64_t)width * 3, (cuuint64_t)c->frame_stride}; // bytes, of dimensions 1 and 2
        const cuuint32_t box[3] = {TILE_COLS * 3, 8, 1}, estr[3] = {1, 1, 1};
        const CUresult cr = ((encode_fn)fn)(&c->fbmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d_frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            g_create_error = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)cr);
            drr_ctx_destroy(c);
            return DRR_E_CUDA;
        }
        c->have_fbmap = true;
    }
    if ((e = cudaMalloc((void **)&c->d_sky_rows, (size_t)height)) != cudaSuccess) return bail("cudaMalloc(sky rows)", e);
    if ((e = launch_sky_rows(c->d_sky_rows, height, c->stream)) != cudaSuccess) return bail("sky rows kernel", e);
    if ((e = cudaMemset(c->d_frames, 0, c->frame_stride * (uint64_t)max_views)) != cudaSuccess) return bail("cudaMemset", e);
    if ((e = cudaMemset(c->d_crc, 0, sizeof(uint64_t) * (size_t)max_views)) != cudaSuccess) return bail("cudaMemset", e);
    for (auto &ev : c->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&c->cstream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate(copy)", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_lists_free, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate(2)", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_stream2, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    *out = c;
    return DRR_OK;
}

void drr_ctx_destroy(drr_ctx *ctx) {
    if (!ctx) return;
    if (ctx->host_only) {
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->cstream) cudaStreamSynchronize(ctx->cstream);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->chunk_ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->prof_ev) cudaEventDestroy(ev);
    if (ctx->ev_lists_free) cudaEventDestroy(ctx->ev_lists_free);
    if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
    if (ctx->cstream) cudaStreamDestroy(ctx->cstream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2), cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_stream2) cudaEventDestroy(ctx->ev_stream2);
    if (ctx->d_frames) cudaFree(ctx->d_frames);
    if (ctx->d_crc) cudaFree(ctx->d_crc);
    if (ctx->d_sky_rows) cudaFree(ctx->d_sky_rows);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int drr_set_stream(drr_ctx *ctx, void *cuda_stream) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
    ctx->stream = (cudaStream_t)cuda_stream;
    if (!cuda_stream) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return DRR_OK;
}
void *drr_get_stream(drr_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int drr_set_knob(drr_ctx *ctx, const char *name, int value) {
    CTX_CHECK(ctx);
    size_t n;
    const Knobs::Name *t = Knobs::table(&n);
    for (size_t i = 0; name && i < n; i++)
        if (!strcmp(name, t[i].name)) {
            ctx->knobs.*(t[i].field) = value;
            return DRR_OK;
        }
    return fail(ctx, DRR_E_INVALID, "drr_set_knob: unknown knob");
}

// ---- assets ------------------------------------------------------------------------------------------------------
int drr_upload_palette(drr_ctx *ctx, const uint8_t rgb[768]) {
    CTX_CHECK(ctx);
    if (!rgb) return fail(ctx, DRR_E_INVALID, "drr_upload_palette: null");
    for (int i = 0; i < 256; i++) {
        const uint32_t r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
        const uint32_t packed = r | (g << 8) | (b << 16);
        float w;
        memcpy(&w, &packed, 4);
        ctx->pal[i] = make_float4((float)r, (float)g, (float)b, w); // `color.r as f32` (bitmap_render.rs:204)
    }
    for (int i = 0; i < 257; i++) { // entry 256 backs the None texel (its colour is never stored)
        const float4 p = ctx->pal[std::min(i, 255)];
        uint32_t r, g, b, w;
        memcpy(&r, &p.x, 4);
        memcpy(&g, &p.y, 4);
        memcpy(&b, &p.z, 4);
        memcpy(&w, &p.w, 4);
        ctx->pal_image[2 * i] = (r >> 16) | (g & 0xffff0000u); // 0..255 as f32 has 16 zero low mantissa bits: bf16 is exact
        ctx->pal_image[2 * i + 1] = b;
        ctx->pal_image[257 * 2 + i] = w;
    }
    ctx->have_pal = true;
    ctx->assets_dirty = true;
    return DRR_OK;
}

int drr_upload_bitmap(drr_ctx *ctx, int id, int w, int h, const int16_t *texels) {
    CTX_CHECK(ctx);
    if (!texels || w <= 0 || h <= 0 || w > 32767 || h > 32767) return fail(ctx, DRR_E_INVALID, "drr_upload_bitmap: bad size (the reference divides by width and height)");
    if (ctx->bitmap_slot.count(id)) return fail(ctx, DRR_E_INVALID, "drr_upload_bitmap: id already uploaded");
    // pool layout: column-major [x][y] with column pitch = next pow2 >= h (one screen column walks ONE texture column, i.e. a
    // contiguous run of texels); values are palette BYTE offsets (index * 16; 256 * 16 = None)
    uint32_t pitch = 1;
    while (pitch < (uint32_t)h) pitch <<= 1;
    BitmapRec r;
    r.base = (uint32_t)ctx->texel_pool.size();
    r.w = (int16_t)w;
    r.h = (int16_t)h;
    r.opaque = 1;
    for (size_t i = 0; i < (size_t)w * h; i++)
        if (texels[i] < -1 || texels[i] > 255) return fail(ctx, DRR_E_INVALID, "drr_upload_bitmap: texel outside -1..255");
    const uint32_t pal0 = ctx->pal_base + SM_PAL; // pool values: where a tile CTA finds the palette entry in its shared memory
    ctx->texel_pool.resize(ctx->texel_pool.size() + (size_t)pitch * w, (uint16_t)(pal0 + TEXEL_NONE_INDEX * PAL_ENTRY));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int16_t t = texels[(size_t)y * w + x];
            if (t < 0) r.opaque = 0;
            ctx->texel_pool[r.base + (size_t)x * pitch + y] = (uint16_t)(pal0 + (t < 0 ? TEXEL_NONE_INDEX : (uint32_t)t) * PAL_ENTRY);
        }
    ctx->bitmap_slot[id] = (int)ctx->bitmaps.size();
    ctx->bitmaps.push_back(r);
    ctx->assets_dirty = true;
    ctx->fes.have_map = false; // a front-end map holds resolved slots: upload it again after the assets
    return DRR_OK;
}

int drr_upload_flat(drr_ctx *ctx, int id, const uint8_t px[4096]) {
    CTX_CHECK(ctx);
    if (!px) return fail(ctx, DRR_E_INVALID, "drr_upload_flat: null");
    if (ctx->flat_slot.count(id)) return fail(ctx, DRR_E_INVALID, "drr_upload_flat: id already uploaded");
    if (ctx->flat_slot.size() >= 32767) return fail(ctx, DRR_E_INVALID, "drr_upload_flat: too many flats");
    ctx->flat_slot[id] = (int)(ctx->flat_pool.size() / 4096);
    ctx->flat_pool.insert(ctx->flat_pool.end(), px, px + 4096);
    ctx->assets_dirty = true;
    ctx->fes.have_map = false;
    return DRR_OK;
}

int drr_set_sky(drr_ctx *ctx, int bitmap_id) {
    CTX_CHECK(ctx);
    auto it = ctx->bitmap_slot.find(bitmap_id);
    if (it == ctx->bitmap_slot.end()) return fail(ctx, DRR_E_ASSET, "drr_set_sky: unknown bitmap id");
    const BitmapRec &r = ctx->bitmaps[it->second];
    // draw_sky hard-codes 256x128 (visplanes.rs:49-50); a smaller texture would index out of bounds in the reference
    if (r.w != 256 || r.h != 128) return fail(ctx, DRR_E_ASSET, "drr_set_sky: sky bitmap must be 256x128");
    ctx->sky_slot = it->second;
    return DRR_OK; // (the front-end reads the sky kind at emit time: no need to drop its map)
}

static int upload_assets(drr_ctx *ctx) {
    if (!ctx->assets_dirty) return DRR_OK;
    if (!ctx->have_pal) return fail(ctx, DRR_E_ASSET, "palette not uploaded");
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, ctx->d_pal_image.reserve(SM_PAL_BYTES / 4));
    CU(ctx, cudaMemcpy(ctx->d_pal_image.p, ctx->pal_image, sizeof(ctx->pal_image), cudaMemcpyHostToDevice));
    CU(ctx, ctx->d_texels.reserve(std::max<size_t>(ctx->texel_pool.size(), 1)));
    if (!ctx->texel_pool.empty())
        CU(ctx, cudaMemcpy(ctx->d_texels.p, ctx->texel_pool.data(), ctx->texel_pool.size() * 2, cudaMemcpyHostToDevice));
    CU(ctx, ctx->d_flats.reserve(std::max<size_t>(ctx->flat_pool.size(), 1)));
    if (!ctx->flat_pool.empty()) CU(ctx, cudaMemcpy(ctx->d_flats.p, ctx->flat_pool.data(), ctx->flat_pool.size(), cudaMemcpyHostToDevice));
    CU(ctx, ctx->d_bitmaps.reserve(std::max<size_t>(ctx->bitmaps.size(), 1)));
    if (!ctx->bitmaps.empty())
        CU(ctx, cudaMemcpy(ctx->d_bitmaps.p, ctx->bitmaps.data(), ctx->bitmaps.size() * sizeof(BitmapRec), cudaMemcpyHostToDevice));
    ctx->assets_dirty = false;
    return DRR_OK;
}

// ---- recording ---------------------------------------------------------------------------------------------------
// Recording is appending: the lists go to the device exactly as emitted (SURVEY 8d's algorithmic bytes plus indices);
// turning per-op lists into per-column lists ("column binning") is the bin kernel's job (drr_tile.cu).
// The same code records into a context's own lists (drr_frame_begin ...) and into a recorder's (drr_recorder_frame_begin
// ...): `L` is where the lists go, `ctx` supplies the asset tables and the screen size, `err` receives the message.
struct drr_recorder {
    Lists lists; // plain memory (set_pinned(false))
    drr_ctx *ctx = nullptr;
    std::string err;
};

static int rfail(std::string &err, int code, const char *msg) {
    err = msg;
    return code;
}

static bool push_frame_bases(Lists &L) {
    L.frame_seg_base.push_back((uint32_t)L.segs.n);
    L.frame_col_base.push_back((uint32_t)L.cols.n);
    L.frame_plane_base.push_back((uint32_t)L.planes.n);
    L.frame_parr_base.push_back((uint32_t)L.parr.n);
    return L.frame_op_base.push((uint32_t)L.ops.n) && L.frame_rec_base.push((uint32_t)L.rec_cap);
}

static int rec_frame_begin(Lists &L, const drr_ctx *ctx, std::string &err, int view_idx, const drr_view *view) {
    if (L.in_frame) return rfail(err, DRR_E_STATE, "frame_begin: previous frame not ended");
    if (!view || view_idx < 0 || view_idx >= ctx->max_views) return rfail(err, DRR_E_INVALID, "frame_begin: bad view index");
    if (L.frame_op_base.n == 0 && !push_frame_bases(L)) return rfail(err, DRR_E_NOMEM, "alloc");
    View v{view->pos_x, view->pos_y, view->floor_height, view->angle, view->cos_angle, view->sin_angle};
    if (!L.views.push(v) || !L.frame_slot.push((uint32_t)view_idx)) return rfail(err, DRR_E_NOMEM, "alloc");
    L.ops_n0 = L.ops.n;
    L.segs_n0 = L.segs.n;
    L.cols_n0 = L.cols.n;
    L.planes_n0 = L.planes.n;
    L.parr_n0 = L.parr.n;
    L.rec0 = L.rec_count;
    L.cap0 = L.rec_cap;
    L.stats0 = L.stats;
    L.cur_slot = view_idx;
    L.in_frame = true;
    return DRR_OK;
}

static int rec_frame_abort(Lists &L, std::string &err) {
    if (!L.in_frame) return rfail(err, DRR_E_STATE, "frame_abort outside a frame");
    L.in_frame = false;
    L.views.n--;
    L.frame_slot.n--;
    L.ops.n = L.ops_n0;
    L.segs.n = L.segs_n0;
    L.cols.n = L.cols_n0;
    L.planes.n = L.planes_n0;
    L.parr.n = L.parr_n0;
    L.rec_count = L.rec0;
    L.rec_cap = L.cap0;
    const uint64_t launches = L.stats.kernel_launches;
    L.stats = L.stats0;
    L.stats.kernel_launches = launches;
    return DRR_OK;
}

static int rec_emit_columns(Lists &L, const drr_ctx *ctx, std::string &err, const drr_seg_hdr *hdr, const drr_col *cols, int n) {
    if (!L.in_frame) return rfail(err, DRR_E_STATE, "emit_columns outside a frame");
    if (!hdr || n < 0 || (n > 0 && !cols)) return rfail(err, DRR_E_INVALID, "emit_columns: null");
    auto it = ctx->bitmap_slot.find(hdr->bitmap_id);
    if (it == ctx->bitmap_slot.end()) return rfail(err, DRR_E_ASSET, "emit_columns: unknown bitmap id");
    SegRec r;
    r.bitmap_slot = (uint32_t)it->second;
    r.light_level = hdr->light_level;
    r.phase = hdr->phase;
    r.lsx = hdr->line_start_x;
    r.lsy = hdr->line_start_y;
    r.lex = hdr->line_end_x;
    r.ley = hdr->line_end_y;
    r.start_offset = hdr->start_offset;
    r.start_x = hdr->start_x;
    r.end_x = hdr->end_x;
    r.bottom_height = hdr->bottom_height;
    r.top_height = hdr->top_height;
    r.offset_x = hdr->offset_x;
    r.offset_y = hdr->offset_y;
    r.tex_base = ctx->bitmaps[it->second].base;
    r.tex_w = ctx->bitmaps[it->second].w;
    r.tex_h = ctx->bitmaps[it->second].h;
    r.tex_opaque = ctx->bitmaps[it->second].opaque;
    r.pad[0] = r.pad[1] = 0;
    static_assert(sizeof(drr_col) == sizeof(ColRec), "drr_col layout");
    if (!L.cols.reserve(L.cols.n + (size_t)n)) return rfail(err, DRR_E_NOMEM, "alloc");
    const int H = ctx->H, W = ctx->W;
    // the bin kernel finds a seg's record for screen column x by index, which needs x strictly increasing within one
    // SegRec: the reference emits x = start_x .. end_x in order; anything else is split into increasing runs
    for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && cols[j].x > cols[j - 1].x) ++j;
        r.cols_first = (uint32_t)L.cols.n;
        r.n = (uint32_t)(j - i);
        r.x0 = cols[i].x;
        r.x1 = cols[j - 1].x;
        memcpy(L.cols.p + L.cols.n, cols + i, sizeof(ColRec) * (size_t)(j - i));
        L.cols.n += (size_t)(j - i);
        for (int k = i; k < j; k++) {
            const drr_col &c = cols[k];
            // Pixels::set ignores x >= W and y > H (pixels.rs:23); negative values become huge usize and are ignored too
            if (c.x < 0 || c.x >= W) continue;
            if (std::max<int>(c.clipped_top_y, 0) <= std::min<int>(c.clipped_bottom_y, H - 1)) L.rec_count++;
        }
        // the bin kernel reserves one record slot per screen column inside the run's x range (its records may be sparser)
        L.rec_cap += (uint64_t)std::max(0, std::min<int>(r.x1, W - 1) - std::max<int>(r.x0, 0) + 1);
        if (!L.ops.push((uint32_t)L.segs.n) || !L.segs.push(r)) return rfail(err, DRR_E_NOMEM, "alloc");
        i = j;
    }
    L.stats.seg_headers++;
    L.stats.column_records += (uint64_t)n;
    return DRR_OK;
}

static int rec_emit_visplane(Lists &L, const drr_ctx *ctx, std::string &err, const drr_visplane_hdr *hdr, const int16_t *top, const int16_t *bottom) {
    if (!L.in_frame) return rfail(err, DRR_E_STATE, "emit_visplane outside a frame");
    if (!hdr || !top || !bottom) return rfail(err, DRR_E_INVALID, "emit_visplane: null");
    const int W = ctx->W, H = ctx->H;
    // the reference indexes [i16; SCREEN_WIDTH] arrays with x (visplanes.rs:61,95): out-of-range x panics there
    if (hdr->left < 0 || hdr->right >= W) return rfail(err, DRR_E_INVALID, "emit_visplane: left/right outside the screen");
    PlaneRec p;
    if (hdr->flat_id == DRR_FLAT_SKY) {
        if (ctx->sky_slot < 0) return rfail(err, DRR_E_ASSET, "emit_visplane: sky not set");
        p.flat_slot = -1;
        p.kind = (int16_t)(ctx->bitmaps[ctx->sky_slot].opaque ? KIND_SKY : KIND_SKY_HOLES);
    } else {
        auto it = ctx->flat_slot.find(hdr->flat_id);
        if (it == ctx->flat_slot.end()) return rfail(err, DRR_E_ASSET, "emit_visplane: unknown flat id");
        p.flat_slot = (int16_t)it->second;
        p.kind = (int16_t)KIND_FLAT;
    }
    p.height = hdr->height;
    p.light_level = hdr->light_level;
    p.left = hdr->left;
    p.right = hdr->right;
    p.arr_first = (uint32_t)L.parr.n;
    const int ncols = hdr->right >= hdr->left ? hdr->right - hdr->left + 1 : 0;
    if (!L.parr.reserve(L.parr.n + (size_t)ncols)) return rfail(err, DRR_E_NOMEM, "alloc");
    for (int i = 0; i < ncols; i++) {
        L.parr.p[L.parr.n++] = (uint32_t)(uint16_t)top[i] | ((uint32_t)(uint16_t)bottom[i] << 16); // unclamped, as stored (quirk Q3)
        const int t = std::max<int>(top[i], 0);            // visplanes.rs:61 / :95
        const int b = std::min<int>(bottom[i], H - 1);     // :62 / :96
        if (p.kind == (int16_t)KIND_FLAT && (int16_t)(b - t) <= 1) continue; // :99-101 (not applied to sky)
        if (t <= b) L.rec_count++;
    }
    L.rec_cap += (uint64_t)ncols;
    if (ncols > 0 && (!L.ops.push(0x80000000u | (uint32_t)L.planes.n) || !L.planes.push(p))) return rfail(err, DRR_E_NOMEM, "alloc");
    L.stats.visplanes++;
    L.stats.visplane_columns += (uint64_t)ncols;
    return DRR_OK;
}

static int rec_frame_end(Lists &L, std::string &err) {
    if (!L.in_frame) return rfail(err, DRR_E_STATE, "frame_end outside a frame");
    if (L.rec_cap > 0xffffffffull || L.cols.n > 0xffffffffull || L.parr.n > 0xffffffffull) {
        rec_frame_abort(L, err);
        return rfail(err, DRR_E_INVALID, "batch too large: more than 2^32 column records");
    }
    L.in_frame = false;
    if (!push_frame_bases(L)) return rfail(err, DRR_E_NOMEM, "alloc");
    L.stats.frames++;
    return DRR_OK;
}

int drr_reset(drr_ctx *ctx) {
    CTX_DEV(ctx);
    if (ctx->in_frame) return fail(ctx, DRR_E_STATE, "drr_reset inside a frame");
    if (int rc = lists_writable(ctx)) return rc;
    ctx->clear_lists();
    ctx->device_lists = false;
    ctx->fes.slab_draw = false;
    ctx->t_spans.clear();
    ctx->t_colidx.clear();
    ctx->uploaded_frames = 0;
    std::fill(ctx->slot_to_frame.begin(), ctx->slot_to_frame.end(), -1);
    const uint64_t launches = ctx->stats.kernel_launches;
    ctx->stats = drr_stats{};
    ctx->stats.kernel_launches = launches;
    return DRR_OK;
}

int drr_frame_begin(drr_ctx *ctx, int view_idx, const drr_view *view) {
    CTX_DEV(ctx);
    if (ctx->device_lists) return fail(ctx, DRR_E_STATE, "drr_frame_begin: the batch was written by drr_fe_emit_views (call drr_reset first)");
    if (int rc = lists_writable(ctx)) return rc;
    if (view_idx >= 0 && view_idx < ctx->max_views && ctx->slot_to_frame[view_idx] >= 0 && !ctx->in_frame)
        return fail(ctx, DRR_E_INVALID, "drr_frame_begin: view index already recorded since drr_reset");
    const int rc = rec_frame_begin(*ctx, ctx, ctx->err, view_idx, view);
    if (rc) return rc;
    ctx->slot_to_frame[view_idx] = (int)ctx->views.n - 1;
    ctx->t_spans.clear();
    ctx->t_colidx.clear();
    return DRR_OK;
}

int drr_frame_abort(drr_ctx *ctx) {
    CTX_CHECK(ctx);
    const int rc = rec_frame_abort(*ctx, ctx->err);
    if (rc == DRR_OK) ctx->slot_to_frame[ctx->cur_slot] = -1;
    return rc;
}

int drr_emit_columns(drr_ctx *ctx, const drr_seg_hdr *hdr, const drr_col *cols, int n) {
    CTX_CHECK(ctx);
    return rec_emit_columns(*ctx, ctx, ctx->err, hdr, cols, n);
}

int drr_emit_visplane(drr_ctx *ctx, const drr_visplane_hdr *hdr, const int16_t *top, const int16_t *bottom) {
    CTX_CHECK(ctx);
    return rec_emit_visplane(*ctx, ctx, ctx->err, hdr, top, bottom);
}

int drr_frame_end(drr_ctx *ctx) {
    CTX_CHECK(ctx);
    const int slot = ctx->cur_slot;
    const int rc = rec_frame_end(*ctx, ctx->err);
    if (rc != DRR_OK && !ctx->in_frame && slot >= 0 && slot < ctx->max_views && ctx->slot_to_frame[slot] == (int)ctx->views.n)
        ctx->slot_to_frame[slot] = -1; // the frame was dropped (batch too large)
    return rc;
}

// ---- recorders: the same recording, off-context (one per worker thread), appended to the context afterwards ------------
int drr_recorder_create(drr_ctx *ctx, drr_recorder **out) {
    if (!ctx || !out) return DRR_E_INVALID;
    drr_recorder *r = new drr_recorder();
    r->ctx = ctx;
    r->lists.set_pinned(false);
    *out = r;
    return DRR_OK;
}
void drr_recorder_destroy(drr_recorder *rec) { delete rec; }
const char *drr_recorder_last_error(const drr_recorder *rec) { return rec ? rec->err.c_str() : ""; }
int drr_recorder_frame_begin(drr_recorder *rec, int view_idx, const drr_view *view) {
    if (!rec) return DRR_E_INVALID;
    return rec_frame_begin(rec->lists, rec->ctx, rec->err, view_idx, view);
}
int drr_recorder_emit_columns(drr_recorder *rec, const drr_seg_hdr *hdr, const drr_col *cols, int n) {
    if (!rec) return DRR_E_INVALID;
    return rec_emit_columns(rec->lists, rec->ctx, rec->err, hdr, cols, n);
}
int drr_recorder_emit_visplane(drr_recorder *rec, const drr_visplane_hdr *hdr, const int16_t *top, const int16_t *bottom) {
    if (!rec) return DRR_E_INVALID;
    return rec_emit_visplane(rec->lists, rec->ctx, rec->err, hdr, top, bottom);
}
int drr_recorder_frame_end(drr_recorder *rec) {
    if (!rec) return DRR_E_INVALID;
    return rec_frame_end(rec->lists, rec->err);
}
int drr_recorder_frame_abort(drr_recorder *rec) {
    if (!rec) return DRR_E_INVALID;
    return rec_frame_abort(rec->lists, rec->err);
}

// Move every frame of `rec` to the end of the context's lists (indices re-based), then clear the recorder.
int drr_append(drr_ctx *ctx, drr_recorder *rec) {
    if (ctx && ctx->device_lists) return fail(ctx, DRR_E_STATE, "drr_append: the batch was written by drr_fe_emit_views (call drr_reset first)");
    CTX_DEV(ctx);
    if (!rec || rec->ctx != ctx) return fail(ctx, DRR_E_INVALID, "drr_append: recorder of another context");
    if (int rc = lists_writable(ctx)) return rc;
    Lists &R = rec->lists;
    if (ctx->in_frame || R.in_frame) return fail(ctx, DRR_E_STATE, "drr_append inside a frame");
    const size_t nf = R.views.n;
    if (nf == 0) return DRR_OK;
    for (size_t i = 0; i < nf; i++) { // every view index must be free in the context and unique in the recorder
        const uint32_t slot = R.frame_slot.p[i];
        if (ctx->slot_to_frame[slot] >= 0) {
            for (size_t k = 0; k < i; k++) ctx->slot_to_frame[R.frame_slot.p[k]] = -1;
            return fail(ctx, DRR_E_INVALID, "drr_append: view index already recorded since drr_reset");
        }
        ctx->slot_to_frame[slot] = (int)(ctx->views.n + i);
    }
    auto undo = [&]() {
        for (size_t i = 0; i < nf; i++) ctx->slot_to_frame[R.frame_slot.p[i]] = -1;
        return fail(ctx, DRR_E_NOMEM, "pinned alloc");
    };
    if (ctx->rec_cap + R.rec_cap > 0xffffffffull || ctx->cols.n + R.cols.n > 0xffffffffull || ctx->parr.n + R.parr.n > 0xffffffffull) {
        for (size_t i = 0; i < nf; i++) ctx->slot_to_frame[R.frame_slot.p[i]] = -1;
        return fail(ctx, DRR_E_INVALID, "batch too large: more than 2^32 column records");
    }
    if (ctx->frame_op_base.n == 0 && !push_frame_bases(*ctx)) return undo();
    const uint32_t op_off = (uint32_t)ctx->ops.n, seg_off = (uint32_t)ctx->segs.n, col_off = (uint32_t)ctx->cols.n;
    const uint32_t plane_off = (uint32_t)ctx->planes.n, parr_off = (uint32_t)ctx->parr.n, cap_off = (uint32_t)ctx->rec_cap;
    if (!ctx->views.reserve(ctx->views.n + nf) || !ctx->frame_slot.reserve(ctx->frame_slot.n + nf) || !ctx->ops.reserve(ctx->ops.n + R.ops.n) ||
        !ctx->segs.reserve(ctx->segs.n + R.segs.n) || !ctx->cols.reserve(ctx->cols.n + R.cols.n) || !ctx->planes.reserve(ctx->planes.n + R.planes.n) ||
        !ctx->parr.reserve(ctx->parr.n + R.parr.n) || !ctx->frame_op_base.reserve(ctx->frame_op_base.n + nf) ||
        !ctx->frame_rec_base.reserve(ctx->frame_rec_base.n + nf))
        return undo();
    memcpy(ctx->views.p + ctx->views.n, R.views.p, nf * sizeof(View));
    ctx->views.n += nf;
    memcpy(ctx->frame_slot.p + ctx->frame_slot.n, R.frame_slot.p, nf * 4);
    ctx->frame_slot.n += nf;
    for (size_t i = 0; i < R.ops.n; i++) {
        const uint32_t op = R.ops.p[i];
        ctx->ops.p[ctx->ops.n++] = (op & 0x80000000u) ? (0x80000000u | ((op & 0x7fffffffu) + plane_off)) : op + seg_off;
    }
    for (size_t i = 0; i < R.segs.n; i++) {
        SegRec g = R.segs.p[i];
        g.cols_first += col_off;
        ctx->segs.p[ctx->segs.n++] = g;
    }
    memcpy(ctx->cols.p + ctx->cols.n, R.cols.p, R.cols.n * sizeof(ColRec));
    ctx->cols.n += R.cols.n;
    for (size_t i = 0; i < R.planes.n; i++) {
        PlaneRec q = R.planes.p[i];
        q.arr_first += parr_off;
        ctx->planes.p[ctx->planes.n++] = q;
    }
    memcpy(ctx->parr.p + ctx->parr.n, R.parr.p, R.parr.n * 4);
    ctx->parr.n += R.parr.n;
    for (size_t i = 1; i <= nf; i++) { // the recorder's bases start at 0
        ctx->frame_op_base.p[ctx->frame_op_base.n++] = R.frame_op_base.p[i] + op_off;
        ctx->frame_rec_base.p[ctx->frame_rec_base.n++] = R.frame_rec_base.p[i] + cap_off;
        ctx->frame_seg_base.push_back(R.frame_seg_base[i] + seg_off);
        ctx->frame_col_base.push_back(R.frame_col_base[i] + col_off);
        ctx->frame_plane_base.push_back(R.frame_plane_base[i] + plane_off);
        ctx->frame_parr_base.push_back(R.frame_parr_base[i] + parr_off);
    }
    ctx->rec_count += R.rec_count;
    ctx->rec_cap += R.rec_cap;
    ctx->stats.frames += R.stats.frames;
    ctx->stats.seg_headers += R.stats.seg_headers;
    ctx->stats.column_records += R.stats.column_records;
    ctx->stats.visplanes += R.stats.visplanes;
    ctx->stats.visplane_columns += R.stats.visplane_columns;
    ctx->t_spans.clear();
    ctx->t_colidx.clear();
    R.clear_lists();
    R.stats = drr_stats{};
    return DRR_OK;
}

// ---- execution ---------------------------------------------------------------------------------------------------
static int reserve_device_lists(drr_ctx *ctx) {
    const size_t nf = ctx->views.n;
    CU(ctx, ctx->d_views.reserve(nf));
    CU(ctx, ctx->d_ops.reserve(std::max<size_t>(ctx->ops.n, 1)));
    CU(ctx, ctx->d_frame_op_base.reserve(nf + 1));
    CU(ctx, ctx->d_frame_rec_base.reserve(nf + 1));
    CU(ctx, ctx->d_frame_slot.reserve(nf));
    CU(ctx, ctx->d_frame_cursor.reserve(nf));
    CU(ctx, ctx->d_segs.reserve(std::max<size_t>(ctx->segs.n, 1)));
    CU(ctx, ctx->d_cols.reserve(std::max<size_t>(ctx->cols.n, 1) * 5));
    CU(ctx, ctx->d_planes.reserve(std::max<size_t>(ctx->planes.n, 1)));
    CU(ctx, ctx->d_parr.reserve(std::max<size_t>(ctx->parr.n, 1)));
    int nbands, band_rows;
    tile_bands(ctx->H, ctx->knobs.tile_max_rows, &nbands, &band_rows);
    const size_t nlists = nbands <= MAX_LIST_BANDS ? (size_t)nbands : 1; // one span list per (column, row band)
    if (ctx->rec_cap * nlists > 0xffffffffull) return fail(ctx, DRR_E_INVALID, "batch too large: more than 2^32 record slots");
    CU(ctx, ctx->d_colidx.reserve(nf * (size_t)ctx->W * nlists));
    CU(ctx, ctx->d_tparams.reserve(std::max<size_t>(ctx->rec_cap, 1) * 4 * nlists));
    return DRR_OK;
}

// H2D of the per-frame tables (small) -- always whole
static int upload_tables(drr_ctx *ctx, cudaStream_t st) {
    const size_t nf = ctx->views.n;
    CU(ctx, cudaMemcpyAsync(ctx->d_views.p, ctx->views.p, nf * sizeof(View), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_op_base.p, ctx->frame_op_base.p, (nf + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_rec_base.p, ctx->frame_rec_base.p, (nf + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_slot.p, ctx->frame_slot.p, nf * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return DRR_OK;
}

// H2D of the draw lists of frames [f0, f1) (contiguous slices: frames are recorded one after the other)
static int upload_frames(drr_ctx *ctx, size_t f0, size_t f1, cudaStream_t st) {
    auto slice = [&](void *dst, const void *src, size_t elem, size_t lo, size_t hi) -> cudaError_t {
        if (hi <= lo) return cudaSuccess;
        return cudaMemcpyAsync((char *)dst + lo * elem, (const char *)src + lo * elem, (hi - lo) * elem, cudaMemcpyHostToDevice, st);
    };
    CU(ctx, slice(ctx->d_ops.p, ctx->ops.p, 4, ctx->frame_op_base.p[f0], ctx->frame_op_base.p[f1]));
    CU(ctx, slice(ctx->d_segs.p, ctx->segs.p, sizeof(SegRec), ctx->frame_seg_base[f0], ctx->frame_seg_base[f1]));
    CU(ctx, slice(ctx->d_cols.p, ctx->cols.p, sizeof(ColRec), ctx->frame_col_base[f0], ctx->frame_col_base[f1]));
    CU(ctx, slice(ctx->d_planes.p, ctx->planes.p, sizeof(PlaneRec), ctx->frame_plane_base[f0], ctx->frame_plane_base[f1]));
    CU(ctx, slice(ctx->d_parr.p, ctx->parr.p, 4, ctx->frame_parr_base[f0], ctx->frame_parr_base[f1]));
    return DRR_OK;
}

static uint64_t list_bytes(const drr_ctx *ctx) {
    return ctx->views.n * sizeof(View) + ctx->ops.n * 4 + ctx->frame_op_base.n * 4 + ctx->frame_rec_base.n * 4 + ctx->frame_slot.n * 4 +
           ctx->segs.n * sizeof(SegRec) + ctx->cols.n * sizeof(ColRec) + ctx->planes.n * sizeof(PlaneRec) + ctx->parr.n * 4;
}

int drr_upload_lists(drr_ctx *ctx) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context: libdrr has no CPU draw path");
    if (ctx->in_frame) return fail(ctx, DRR_E_STATE, "drr_upload_lists inside a frame");
    if (ctx->device_lists) return DRR_OK; // drr_fe_emit_views wrote the lists on the device: nothing to upload
    int rc = upload_assets(ctx);
    if (rc) return rc;
    const size_t nf = ctx->views.n;
    ctx->uploaded_frames = 0;
    if (nf == 0) return DRR_OK;
    if ((rc = reserve_device_lists(ctx))) return rc;
    if ((rc = upload_tables(ctx, ctx->stream))) return rc;
    if ((rc = upload_frames(ctx, 0, nf, ctx->stream))) return rc;
    CU(ctx, cudaEventRecord(ctx->ev_h2d, ctx->stream));
    ctx->h2d_pending = true;
    ctx->uploaded_frames = nf;
    return DRR_OK;
}

static int make_args(drr_ctx *ctx, DrawArgs &a, size_t nframes) {
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context: libdrr has no CPU draw path");
    if (nframes == 0) return fail(ctx, DRR_E_STATE, "nothing uploaded (call drr_upload_lists)");
    a.W = ctx->W;
    a.H = ctx->H;
    a.nframes = (int)nframes;
    tile_bands(ctx->H, ctx->knobs.tile_max_rows, &a.nbands, &a.band_rows);
    a.CFX = ctx->CFX;
    a.CFY = ctx->CFY;
    a.GCFX = ctx->GCFX;
    a.ASPECT = ctx->ASPECT;
    a.Wf = (float)(uint32_t)ctx->W;
    a.Hf = (float)(uint32_t)ctx->H;
    a.one = 1.0f;
    a.dbg = ctx->knobs.dbg; // timing experiments of a -DDRR_DBG_KNOBS build only
    a.views = ctx->d_views.p;
    a.ops = ctx->d_ops.p;
    a.frame_op_base = ctx->d_frame_op_base.p;
    a.frame_nops = nullptr;
    a.frame_rec_base = ctx->d_frame_rec_base.p;
    a.frame_slot = ctx->d_frame_slot.p;
    a.segs = ctx->d_segs.p;
    a.cols = reinterpret_cast<const ColRec *>(ctx->d_cols.p);
    a.planes = ctx->d_planes.p;
    a.parr = ctx->d_parr.p;
    a.frame_cursor = ctx->d_frame_cursor.p;
    a.colidx = ctx->d_colidx.p;
    a.tparams = ctx->d_tparams.p;
    a.texels = ctx->d_texels.p;
    a.flats = ctx->d_flats.p;
    a.bitmaps = ctx->d_bitmaps.p;
    a.pal_image = ctx->d_pal_image.p;
    a.pal_base = ctx->pal_base;
    a.sky_rows = ctx->d_sky_rows;
    a.sky_base = ctx->sky_slot >= 0 ? ctx->bitmaps[ctx->sky_slot].base : 0;
    a.frames = ctx->d_frames;
    a.frame_stride = ctx->frame_stride;
    a.crc = ctx->d_crc;
    if (ctx->device_lists && ctx->fes.slab_draw) { // the front-end's per-view slabs as they are: every index in them is an index into the slab arrays
        const FeState &S = ctx->fes;
        a.ops = S.sl_last.out.ops;
        a.segs = S.sl_last.out.segs;
        a.cols = S.sl_last.out.cols;
        a.planes = S.sl_last.out.planes;
        a.parr = S.sl_last.out.parr;
        a.frame_op_base = S.d_slab_meta.p;
        a.frame_nops = S.d_slab_meta.p + nframes + 1;
    }
    return DRR_OK;
}

// bin kernel + tile kernel over frames [f0, f0 + n) on the context's stream
static int draw_range(drr_ctx *ctx, const DrawArgs &a, int f0, int n, bool bin, bool tile, bool profile = false, cudaStream_t st = nullptr) {
    if (!st) st = ctx->stream;
    const bool prof = profile && (size_t)(ctx->prof_steps + 1) * 3 <= ctx->prof_ev.size();
    if (prof) CU(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_steps * 3], st));
    if (bin) {
        CU(ctx, launch_bin(a, f0, n, st));
        ctx->stats.kernel_launches++;
    }
    if (prof) CU(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_steps * 3 + 1], st));
    if (tile) {
        int launches = 0;
        CU(ctx, launch_tile(a, ctx->have_fbmap ? &ctx->fbmap : nullptr, f0, n, st, &launches));
        ctx->stats.kernel_launches += (uint64_t)launches;
    }
    if (prof) {
        CU(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_steps * 3 + 2], st));
        ctx->prof_steps++;
    }
    return DRR_OK;
}

int drr_draw(drr_ctx *ctx) {
    CTX_DEV(ctx);
    DrawArgs a;
    int rc = make_args(ctx, a, ctx->uploaded_frames);
    if (rc) return rc;
    CU(ctx, cudaMemsetAsync(ctx->d_crc, 0, sizeof(uint64_t) * (size_t)ctx->max_views, ctx->stream));
    return draw_range(ctx, a, 0, a.nframes, true, true, true);
}

// Per-kernel device times of the drr_draw() calls made since drr_profile_begin, from CUDA events recorded on the
// context's stream around each kernel.
int drr_profile_begin(drr_ctx *ctx, int max_steps) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    if (max_steps < 0 || max_steps > 4096) return fail(ctx, DRR_E_INVALID, "drr_profile_begin: max_steps");
    while (ctx->prof_ev.size() < (size_t)max_steps * 3) {
        cudaEvent_t e;
        CU(ctx, cudaEventCreate(&e));
        ctx->prof_ev.push_back(e);
    }
    ctx->prof_steps = 0;
    return DRR_OK;
}
int drr_profile_end(drr_ctx *ctx, int *steps, float *setup_ms_total, float *march_ms_total) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    float s = 0, m = 0;
    for (int i = 0; i < ctx->prof_steps; i++) {
        float a = 0, b = 0;
        CU(ctx, cudaEventElapsedTime(&a, ctx->prof_ev[i * 3], ctx->prof_ev[i * 3 + 1]));
        CU(ctx, cudaEventElapsedTime(&b, ctx->prof_ev[i * 3 + 1], ctx->prof_ev[i * 3 + 2]));
        s += a;
        m += b;
    }
    if (steps) *steps = ctx->prof_steps;
    if (setup_ms_total) *setup_ms_total = s;
    if (march_ms_total) *march_ms_total = m;
    ctx->prof_steps = 0;
    for (auto e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    return DRR_OK;
}

// drr_submit: upload and draw, pipelined.  The batch is cut into chunks of frames; chunk k+1's lists travel over PCIe
// on the copy stream while chunk k is binned and drawn on the context's stream.
int drr_submit(drr_ctx *ctx) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context: libdrr has no CPU draw path");
    if (ctx->in_frame) return fail(ctx, DRR_E_STATE, "drr_submit inside a frame");
    if (ctx->device_lists) return drr_draw(ctx); // drr_fe_emit_views wrote the lists on the device: nothing to upload
    int rc = upload_assets(ctx);
    if (rc) return rc;
    const size_t nf = ctx->views.n;
    ctx->uploaded_frames = 0;
    if (nf == 0) return DRR_OK;
    if ((rc = reserve_device_lists(ctx))) return rc;
    // ~4 MB of lists per chunk, at most 8 chunks (measured: tools/sweep_submit.sh; more chunks cost more in launches than they hide)
    size_t nchunks = std::min<size_t>({(size_t)8, nf, (size_t)(list_bytes(ctx) / (4u << 20)) + 1});
    if (ctx->knobs.submit_chunks > 0) nchunks = std::min<size_t>(nf, (size_t)ctx->knobs.submit_chunks);
    while (ctx->chunk_ev.size() < nchunks) {
        cudaEvent_t e;
        CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(e);
    }
    DrawArgs a;
    if ((rc = make_args(ctx, a, nf))) return rc;
    // the copies overwrite lists that kernels already queued on the context's stream may still read
    CU(ctx, cudaEventRecord(ctx->ev_lists_free, ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->cstream, ctx->ev_lists_free, 0));
    CU(ctx, cudaMemsetAsync(ctx->d_crc, 0, sizeof(uint64_t) * (size_t)ctx->max_views, ctx->stream));
    const bool two = nchunks > 1 && !ctx->knobs.submit_one_stream;
    if (two) { // stream2's first kernel must come after the checksum memset queued on the context's stream
        CU(ctx, cudaEventRecord(ctx->ev_stream2, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_stream2, 0));
    }
    if ((rc = upload_tables(ctx, ctx->cstream))) return rc;
    // DRR_SUBMIT_TRACE=1: print when each chunk's copy and draw finished (device time since the start of the call)
    const bool trace = ctx->knobs.submit_trace != 0;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(1 + 2 * nchunks);
        for (auto &e : tev) CU(ctx, cudaEventCreate(&e));
        CU(ctx, cudaEventRecord(tev[0], ctx->cstream));
    }
    for (size_t c = 0; c < nchunks; c++) {
        const size_t f0 = nf * c / nchunks, f1 = nf * (c + 1) / nchunks;
        if ((rc = upload_frames(ctx, f0, f1, ctx->cstream))) return rc;
        CU(ctx, cudaEventRecord(ctx->chunk_ev[c], ctx->cstream));
        if (trace) CU(ctx, cudaEventRecord(tev[1 + 2 * c], ctx->cstream));
        cudaStream_t st = (two && (c & 1)) ? ctx->stream2 : ctx->stream;
        CU(ctx, cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0));
        if ((rc = draw_range(ctx, a, (int)f0, (int)(f1 - f0), true, true, false, st))) return rc;
        if (trace) CU(ctx, cudaEventRecord(tev[2 + 2 * c], st));
    }
    CU(ctx, cudaEventRecord(ctx->ev_h2d, ctx->cstream));
    ctx->h2d_pending = true;
    if (two) { // everything is complete when the context's stream is
        CU(ctx, cudaEventRecord(ctx->ev_stream2, ctx->stream2));
        CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_stream2, 0));
    }
    ctx->uploaded_frames = nf;
    if (trace) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t c = 0; c < nchunks; c++) {
            float tc = 0, td = 0;
            cudaEventElapsedTime(&tc, tev[0], tev[1 + 2 * c]);
            cudaEventElapsedTime(&td, tev[0], tev[2 + 2 * c]);
            fprintf(stderr, "drr_submit chunk %zu/%zu: copy done %.3f ms, draw done %.3f ms\n", c, nchunks, tc, td);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return DRR_OK;
}

int drr_sync(drr_ctx *ctx) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return DRR_OK;
}

int drr_read_framebuffer(drr_ctx *ctx, int view_idx, uint8_t *out) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    if (!out || view_idx < 0 || view_idx >= ctx->max_views) return fail(ctx, DRR_E_INVALID, "drr_read_framebuffer: bad view index");
    CU(ctx, cudaMemcpyAsync(out, ctx->d_frames + (uint64_t)view_idx * ctx->frame_stride, (size_t)ctx->W * ctx->H * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return DRR_OK;
}

int drr_read_checksums(drr_ctx *ctx, int first, int count, uint64_t *out) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    if (!out || first < 0 || count < 0 || first + count > ctx->max_views) return fail(ctx, DRR_E_INVALID, "drr_read_checksums: bad range");
    if (count == 0) return DRR_OK;
    if (!ctx->h_crc.reserve((size_t)count)) return fail(ctx, DRR_E_NOMEM, "pinned alloc");
    CU(ctx, cudaMemcpyAsync(ctx->h_crc.p, ctx->d_crc + first, sizeof(uint64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->h_crc.p, sizeof(uint64_t) * (size_t)count);
    return DRR_OK;
}

int drr_read_crc32(drr_ctx *ctx, int first, int count, uint32_t *out) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    if (!out || first < 0 || count < 0 || first + count > ctx->max_views) return fail(ctx, DRR_E_INVALID, "drr_read_crc32: bad range");
    if (count == 0) return DRR_OK;
    const uint64_t nbytes = (uint64_t)ctx->W * ctx->H * 3;
    DevBuf<uint32_t> scratch, result;
    CU(ctx, scratch.reserve(crc32_scratch_words(nbytes, count)));
    CU(ctx, result.reserve((size_t)count));
    CU(ctx, launch_crc32(ctx->d_frames, ctx->frame_stride, nbytes, first, count, scratch.p, result.p, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, result.p, sizeof(uint32_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return DRR_OK;
}

uint64_t drr_checksum_host(const uint8_t *rgb24, uint64_t nbytes) {
    uint64_t acc = 0;
    const uint64_t ngroups = (nbytes + 4 * CK_GROUP - 1) / (4 * CK_GROUP);
    for (uint64_t g = 0; g < ngroups; g++) {
        uint32_t sum = 0;
        for (int j = 0; j < CK_GROUP; j++) {
            const uint64_t i = g * CK_GROUP + j;
            uint32_t w = 0;
            for (int k = 0; k < 4; k++)
                if (i * 4 + k < nbytes) w |= (uint32_t)rgb24[i * 4 + k] << (8 * k);
            sum += w * checksum_word_weight(j);
        }
        acc += checksum_group_term(sum, g);
    }
    return acc;
}

int drr_get_stats(drr_ctx *ctx, drr_stats *out) {
    CTX_CHECK(ctx);
    if (!out) return DRR_E_INVALID;
    drr_stats s = ctx->stats;
    s.spans = ctx->rec_count;
    s.drawlist_bytes_algorithmic = 24 * s.frames + 48 * s.seg_headers + 10 * s.column_records + 12 * s.visplanes + 4 * s.visplane_columns;
    s.device_list_bytes = ctx->device_lists ? ctx->fes.device_list_bytes : list_bytes(ctx);
    *out = s;
    return DRR_OK;
}

int drr_time_draw(drr_ctx *ctx, int iters, float *total_ms, float *setup_ms, float *march_ms) {
    CTX_DEV(ctx);
    if (iters <= 0) return fail(ctx, DRR_E_INVALID, "drr_time_draw: iters");
    DrawArgs a;
    int rc = make_args(ctx, a, ctx->uploaded_frames);
    if (rc) return rc;
    struct { bool s, m; float *out; } legs[3] = {{true, true, total_ms}, {true, false, setup_ms}, {false, true, march_ms}};
    for (auto &leg : legs) {
        if (!leg.out) continue;
        CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        for (int i = 0; i < iters; i++) {
            if (leg.m) CU(ctx, cudaMemsetAsync(ctx->d_crc, 0, sizeof(uint64_t) * (size_t)ctx->max_views, ctx->stream));
            if ((rc = draw_range(ctx, a, 0, a.nframes, leg.s, leg.m))) return rc;
        }
        CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(ctx, cudaEventSynchronize(ctx->ev[1]));
        float ms = 0;
        CU(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        *leg.out = ms / (float)iters;
    }
    return DRR_OK;
}

// ---- device front-end ---------------------------------------------------------------------------
// drr_fe_upload_map flattens nothing itself: the caller (csrc/host/drr_scene.cpp, or a Rust host) hands over the map's
// tables; ids are resolved to slots here.  drr_fe_emit_views = count pass, host offsets, emit pass (drr_frontend.cu).
static const char *fe_detail_message(uint32_t d) {
    switch (d) {
    case fe::FED_CLIP_X: return "Clipped line x < -0.01";
    case fe::FED_UNKNOWN_TEXTURE: return "Unknown texture";
    case fe::FED_NOT_VERTICAL: return "Wall start not vertical";
    case fe::FED_LINE_X: return "Invalid line start/end x";
    case fe::FED_FLAT_MISSING: return "flat lump missing";
    case fe::FED_STACK: return "BSP deeper than the front-end's walk stack";
    case fe::FED_BITMAP_SLOT: return "a wall's bitmap was never uploaded";
    case fe::FED_SKY_UNSET: return "sky visplane but no sky bitmap set";
    case fe::FED_SCRATCH: return "the view outgrew the front-end's working arrays";
    case fe::FED_ROTATION: return "Invalid rotation";
    case fe::FED_MO_X: return "map object column outside the screen";
    default: return "front-end error";
    }
}

int drr_fe_upload_map(drr_ctx *ctx, const drr_fe_map *m) {
    CTX_DEV(ctx);
    if (!m || m->n_nodes <= 0 || m->n_subsectors <= 0 || m->n_segs <= 0 || m->n_linedefs <= 0 || m->n_sidedefs <= 0 || m->n_sectors <= 0 ||
        !m->nodes || !m->subsectors || !m->segs || !m->linedefs || !m->sidedefs || !m->sectors)
        return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: empty table");
    static_assert(sizeof(drr_fe_node) == sizeof(fe::Node) && sizeof(drr_fe_subsector) == sizeof(fe::SubSector) && sizeof(drr_fe_seg) == sizeof(fe::Seg) &&
                      sizeof(drr_fe_linedef) == sizeof(fe::Line) && sizeof(drr_fe_sidedef) == sizeof(fe::Side) && sizeof(drr_fe_sector) == sizeof(fe::Sector) &&
                      sizeof(drr_fe_thing) == sizeof(fe::Thing),
                  "drr_fe_* layouts");
    FeState &S = ctx->fes;
    S.have_map = false;
    // every index the walk follows is checked once here, so the kernel needs no bounds tests
    for (int i = 0; i < m->n_nodes; i++)
        for (int32_t ch : {m->nodes[i].right, m->nodes[i].left})
            // (a node lump stores children before their parent, the root last: anything else could be a cycle)
            if (ch >= i || (ch < 0 && ~ch >= m->n_subsectors)) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: node child out of range or not before its parent");
    for (int i = 0; i < m->n_subsectors; i++)
        if (m->subsectors[i].first_seg < 0 || m->subsectors[i].count < 0 || (int64_t)m->subsectors[i].first_seg + m->subsectors[i].count > m->n_segs)
            return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: subsector seg range");
    for (int i = 0; i < m->n_segs; i++)
        if (m->segs[i].linedef < 0 || m->segs[i].linedef >= m->n_linedefs) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: seg linedef");
    for (int i = 0; i < m->n_linedefs; i++)
        for (int32_t sd : {m->linedefs[i].front, m->linedefs[i].back})
            if (sd < -1 || sd >= m->n_sidedefs) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: linedef sidedef");
    int max_id = -1;
    for (auto &kv : ctx->bitmap_slot) max_id = std::max(max_id, kv.first);
    for (int i = 0; i < m->n_sidedefs; i++) {
        const drr_fe_sidedef &sd = m->sidedefs[i];
        if (sd.sector < 0 || sd.sector >= m->n_sectors) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: sidedef sector");
        for (int32_t t : {sd.upper, sd.lower, sd.middle}) {
            if (t < -2 || t > (1 << 24)) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: sidedef texture id");
            max_id = std::max(max_id, t);
        }
    }
    S.bitmaps.assign((size_t)(max_id + 1), fe::Bitmap{-1, 0u, 0, 0, 0u});
    for (auto &kv : ctx->bitmap_slot)
        if (kv.first >= 0) {
            const BitmapRec &r = ctx->bitmaps[kv.second];
            S.bitmaps[kv.first] = fe::Bitmap{kv.second, r.base, r.w, r.h, r.opaque};
        }
    S.nodes.assign(reinterpret_cast<const fe::Node *>(m->nodes), reinterpret_cast<const fe::Node *>(m->nodes) + m->n_nodes);
    S.ssectors.assign(reinterpret_cast<const fe::SubSector *>(m->subsectors), reinterpret_cast<const fe::SubSector *>(m->subsectors) + m->n_subsectors);
    // The tree seen from below, for the closed-form seg order (drr_frontend.cuh: Frame::run).  Only when the lump is a proper
    // tree: every node but the root and every subsector is the child of exactly one node, and no seg is listed twice.
    {
        const int nn = m->n_nodes, ns = m->n_subsectors;
        S.node_up.assign((size_t)nn, fe::NodeUp{-1, 0, 0});
        S.ss_up.assign((size_t)ns, -2);
        std::vector<int> refs((size_t)nn, 0);
        std::vector<int32_t> total((size_t)nn, 0); // segs under the node
        bool tree = true;
        int64_t nord = 0;
        for (int i = 0; i < ns; i++) nord += m->subsectors[i].count;
        for (int i = 0; i < nn && tree; i++) { // children come before their parents (checked above)
            for (int side = 0; side < 2; side++) {
                const int32_t ch = side ? m->nodes[i].left : m->nodes[i].right;
                const int32_t up = (int32_t)((uint32_t)i << 1) | side;
                int32_t under;
                if (ch < 0) {
                    if (S.ss_up[(size_t)~ch] != -2) tree = false;
                    S.ss_up[(size_t)~ch] = up;
                    under = m->subsectors[~ch].count;
                } else {
                    if (++refs[(size_t)ch] > 1) tree = false;
                    S.node_up[(size_t)ch].up = up;
                    under = total[(size_t)ch];
                }
                (side ? S.node_up[(size_t)i].left : S.node_up[(size_t)i].right) = under;
                total[(size_t)i] += under;
            }
        }
        for (int i = 0; i + 1 < nn; i++) tree = tree && refs[(size_t)i] == 1;
        for (int i = 0; i < ns; i++) tree = tree && S.ss_up[(size_t)i] != -2;
        tree = tree && nord <= m->n_segs && total[(size_t)nn - 1] == nord;
        if (!tree) {
            S.node_up.clear();
            S.ss_up.clear();
        }
        S.nord = (int)std::min<int64_t>(nord, m->n_segs);
    }
    S.segs.assign(reinterpret_cast<const fe::Seg *>(m->segs), reinterpret_cast<const fe::Seg *>(m->segs) + m->n_segs);
    S.lines.assign(reinterpret_cast<const fe::Line *>(m->linedefs), reinterpret_cast<const fe::Line *>(m->linedefs) + m->n_linedefs);
    S.sides.assign(reinterpret_cast<const fe::Side *>(m->sidedefs), reinterpret_cast<const fe::Side *>(m->sidedefs) + m->n_sidedefs);
    S.sectors.assign(reinterpret_cast<const fe::Sector *>(m->sectors), reinterpret_cast<const fe::Sector *>(m->sectors) + m->n_sectors);
    for (fe::Sector &sec : S.sectors)
        for (int16_t *f : {&sec.floor_flat, &sec.ceil_flat}) { // flat id -> flat slot
            if (*f < 0) {
                *f = -2;
                continue;
            }
            auto it = ctx->flat_slot.find(*f);
            *f = it == ctx->flat_slot.end() ? (int16_t)-2 : (int16_t)it->second;
        }
    S.things.clear();
    if (m->n_things < 0 || (m->n_things > 0 && !m->things)) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: things");
    for (int i = 0; i < m->n_things; i++) {
        const drr_fe_thing &t = m->things[i];
        if (t.sector < -1 || t.sector >= m->n_sectors) return fail(ctx, DRR_E_INVALID, "drr_fe_upload_map: thing sector");
        for (int r = 0; r < (t.rotate ? 8 : 1); r++)
            if (t.bitmap[r] < 0 || t.bitmap[r] >= (int)S.bitmaps.size()) return fail(ctx, DRR_E_ASSET, "drr_fe_upload_map: thing bitmap id was never uploaded");
        S.things.push_back(*reinterpret_cast<const fe::Thing *>(&t));
    }
    if (!ctx->host_only) {
        auto up = [&](auto &dev, const auto &host) -> cudaError_t {
            cudaError_t e = dev.reserve(std::max<size_t>(host.size(), 1));
            if (e != cudaSuccess || host.empty()) return e;
            return cudaMemcpyAsync(dev.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, ctx->stream);
        };
        CU(ctx, cudaStreamSynchronize(ctx->stream)); // a previous batch's front-end may still read the old tables
        CU(ctx, up(S.d_nodes, S.nodes));
        CU(ctx, up(S.d_ssectors, S.ssectors));
        CU(ctx, up(S.d_node_up, S.node_up));
        CU(ctx, up(S.d_ss_up, S.ss_up));
        CU(ctx, up(S.d_segs, S.segs));
        CU(ctx, up(S.d_lines, S.lines));
        CU(ctx, up(S.d_sides, S.sides));
        CU(ctx, up(S.d_sectors, S.sectors));
        CU(ctx, up(S.d_bitmaps, S.bitmaps));
        CU(ctx, up(S.d_things, S.things));
        CU(ctx, cudaStreamSynchronize(ctx->stream)); // the host vectors may be reassigned by the next call
    }
    S.have_map = true;
    static uint32_t next_id = 0; // (process-wide: a new context at a recycled address still gets a fresh id)
    S.map_id = ++next_id ? next_id : ++next_id;
    return DRR_OK;
}
uint32_t drr_fe_map_id(drr_ctx *ctx) { return ctx && ctx->fes.have_map ? ctx->fes.map_id : 0u; }

static fe::Map fe_make_map(const drr_ctx *ctx, bool device, int phases) {
    const FeState &S = ctx->fes;
    fe::Map m;
    m.nodes = device ? S.d_nodes.p : S.nodes.data();
    m.ssectors = device ? S.d_ssectors.p : S.ssectors.data();
    m.node_up = S.node_up.empty() ? nullptr : device ? S.d_node_up.p : S.node_up.data();
    m.ss_up = S.ss_up.empty() ? nullptr : device ? S.d_ss_up.p : S.ss_up.data();
    m.nssectors = (int)S.ssectors.size();
    m.nord = S.nord;
    m.side_words = ((int)S.nodes.size() + 31) / 32;
    m.segs = device ? S.d_segs.p : S.segs.data();
    m.lines = device ? S.d_lines.p : S.lines.data();
    m.sides = device ? S.d_sides.p : S.sides.data();
    m.sectors = device ? S.d_sectors.p : S.sectors.data();
    m.bitmaps = device ? S.d_bitmaps.p : S.bitmaps.data();
    m.things = device ? S.d_things.p : S.things.data();
    m.nthings = (int)S.things.size();
    m.nnodes = (int)S.nodes.size();
    m.nsegs = (int)S.segs.size();
    m.W = ctx->W;
    m.H = ctx->H;
    m.ASPECT = ctx->ASPECT;
    m.GCFX = ctx->GCFX;
    m.CFX = ctx->CFX;
    m.CFY = ctx->CFY;
    m.sky_kind = ctx->sky_slot < 0 ? -1 : (int)(ctx->bitmaps[ctx->sky_slot].opaque ? KIND_SKY : KIND_SKY_HOLES);
    m.phases = phases;
    return m;
}

// The whole batch.  Single-pass mode (default): the emit pass writes every view's lists into fixed-size per-view slabs
// and leaves the counts; the host turns the counts into offsets (exclusive scan over the views that got a frame) and
// drr_fe_compact_kernel copies the slabs into the dense arrays drr_bin_kernel reads.  If a view outgrows its slab, or
// DRR_FE_TWO_PASS is set, the batch runs in two-pass mode instead: count pass, offsets, emit pass straight into the dense
// arrays.  Both modes produce the same bytes.  on_host == true runs the SAME per-view code on the CPU into the context's
// host lists (test infrastructure: lets the CPU suite compare it with the host front-end list by list).
static fe::Caps fe_slab_caps(const drr_ctx *ctx) {
    // room per view: E1M1-class frames use ~1.3 W column records and ~2.5 W visplane columns, the 1920x1200 stress map
    // ~7.5 W and ~12.5 W; DRR_FE_SLAB_DIV shrinks the slabs (tests: provoke the fallback)
    uint32_t W = (uint32_t)ctx->W * ctx->fes.slab_boost, div = 1;
    if (ctx->knobs.fe_slab_div > 0) div = (uint32_t)ctx->knobs.fe_slab_div;
    const uint32_t k = ctx->fes.slab_boost; // the record-count capacities grow with the boost as well
    return fe::Caps{std::max(8u, 2048u * k / div), std::max(4u, 1024u * k / div), std::max(32u, std::max(16u * W, 6144u) / div), std::max(4u, 1024u * k / div),
                    std::max(32u, std::max(24u * W, 12288u) / div)};
}

static int fe_emit_views(drr_ctx *ctx, int first_view_idx, const float *xya, int n, int phases, int *status, bool on_host) {
    FeState &S = ctx->fes;
    if (!S.have_map) return fail(ctx, DRR_E_STATE, "drr_fe_emit_views: no map (drr_fe_upload_map)");
    if (n < 0 || (n > 0 && !xya) || first_view_idx < 0 || (int64_t)first_view_idx + n > ctx->max_views || (phases & ~7))
        return fail(ctx, DRR_E_INVALID, "drr_fe_emit_views: bad arguments");
    if (ctx->in_frame || ctx->views.n != 0 || ctx->device_lists) return fail(ctx, DRR_E_STATE, "drr_fe_emit_views: frames already recorded (call drr_reset first)");
    if (n == 0) return DRR_OK;
    if (!on_host && ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context: libdrr has no CPU path");
    const size_t N = (size_t)n, W = (size_t)ctx->W;
    // DRR_FE_TRACE=1: host-side time line of the call on stderr (where the microseconds between the kernels go)
    const bool trace = ctx->knobs.fe_trace != 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (trace) fprintf(stderr, "drr_fe_emit_views %8.1f us  %s\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(), what);
    };
    if (!S.h_views_in.reserve(N) || !S.h_counts.reserve(N) || !S.h_bases.reserve(N)) return fail(ctx, DRR_E_NOMEM, "alloc");
    for (size_t i = 0; i < N; i++) { // host libm per view (the reference's Vertex::rotate calls): ~4 ns per call, 4 calls per view
        const float a = xya[3 * i + 2];
        S.h_views_in.p[i] = fe::ViewIn{xya[3 * i], xya[3 * i + 1], a, cosf(a), sinf(a), cosf(-a), sinf(-a)};
    }
    mark("per-view cos/sin done");
    const fe::Caps nocap{0, 0, 0, 0, 0}, unlimited{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    fe::Caps slab = fe_slab_caps(ctx);
    bool single = !ctx->knobs.fe_two_pass;
    if ((uint64_t)N * std::max({slab.ops, slab.segs, slab.cols, slab.planes, slab.parr}) > 0xffffffffull) single = false; // slab indices are 32-bit
    S.single_pass = false;
    S.count_ms = S.emit_ms = 0.0f;

    // host harness state: one viewpoint at a time, so a single view's scratch; slabs in plain memory
    std::vector<uint8_t> hs_hor;
    std::vector<int16_t> hs_focl, hs_cocl;
    std::vector<uint32_t> hs_rows, hsl_ops, hsl_parr;
    std::vector<int32_t> hs_order;
    std::vector<View> hsl_views;
    std::vector<SegRec> hsl_segs;
    std::vector<ColRec> hsl_cols;
    std::vector<PlaneRec> hsl_planes;
    FeScratch scr{};
    // masked phase: room per view for the remembered parts, their columns and the sprites' (a part has at most W columns;
    // E1M1-class frames remember ~150 parts / ~2.5 W columns, the stress map ~700 / ~12 W)
    const bool masked = (phases & 4) != 0;
    auto knob_u = [](int v, uint32_t dflt) { return v > 0 ? (uint32_t)v : dflt; };
    uint32_t cap_renders = masked ? knob_u(ctx->knobs.fe_cap_renders, 4096u) * S.scratch_boost : 0u;
    uint32_t cap_dsegs = masked ? knob_u(ctx->knobs.fe_cap_dsegs, 2048u) * S.scratch_boost + (uint32_t)S.things.size() : 0u;
    const uint32_t cap_mos = masked ? (uint32_t)S.things.size() + 1u : 0u;
    uint32_t cap_allcols = masked ? std::max<uint32_t>(knob_u(ctx->knobs.fe_cap_allcols_per_w, 48u) * (uint32_t)W, 16384u) * S.scratch_boost : 0u;
    std::vector<fe::RenderRec> hs_renders;
    std::vector<ColRec> hs_allcols;
    std::vector<SegRec> hs_dsegs;
    std::vector<fe::MoRec> hs_mos(cap_mos);
    std::vector<int32_t> hs_mo_order(cap_mos);
    std::vector<int32_t> hs_dseg_part;
    auto masked_scratch = [&]() -> int { // (re)allocate the masked phase's working arrays for the current capacities
        if (on_host) {
            hs_renders.resize(cap_renders);
            hs_allcols.resize(cap_allcols);
            hs_dsegs.resize(cap_dsegs);
            hs_dseg_part.resize(cap_dsegs);
            return DRR_OK;
        }
        if (masked) {
            CU(ctx, S.d_renders.reserve(N * cap_renders));
            CU(ctx, S.d_allcols.reserve(N * cap_allcols * 5));
            CU(ctx, S.d_dsegs.reserve(N * cap_dsegs));
            CU(ctx, S.d_dseg_part.reserve(N * cap_dsegs));
            CU(ctx, S.d_mos.reserve(N * cap_mos));
            CU(ctx, S.d_mo_order.reserve(N * cap_mos));
        }
        scr = FeScratch{S.d_hor.p, S.d_focl.p, S.d_cocl.p, S.d_rows.p, S.d_order.p, S.d_pre.p, S.d_pre_code.p, S.d_renders.p, S.d_allcols.p, S.d_dsegs.p, S.d_mos.p, S.d_mo_order.p,
                        S.d_dseg_part.p, cap_renders, cap_allcols, cap_dsegs, cap_mos};
        return DRR_OK;
    };
    if (on_host) {
        hs_hor.resize(W);
        hs_focl.resize(W);
        hs_cocl.resize(W);
        hs_rows.resize(2 * W);
        hs_order.resize(S.segs.size() + (S.nodes.size() + 31) / 32); // the seg order, then the node side bits
    } else {
        int rc = upload_assets(ctx);
        if (rc) return rc;
        CU(ctx, S.d_views_in.reserve(N));
        CU(ctx, S.d_counts.reserve(N));
        CU(ctx, S.d_bases.reserve(N));
        CU(ctx, S.d_hor.reserve(N * W));
        CU(ctx, S.d_focl.reserve(N * W));
        CU(ctx, S.d_cocl.reserve(N * W));
        CU(ctx, S.d_rows.reserve(N * W * 2));
        CU(ctx, S.d_order.reserve(N * (S.segs.size() + (S.nodes.size() + 31) / 32))); // per view: the seg order, then the node side bits
        CU(ctx, S.d_pre.reserve(N * S.segs.size()));
        CU(ctx, S.d_pre_code.reserve(N * S.segs.size()));
        CU(ctx, cudaMemcpyAsync(S.d_views_in.p, S.h_views_in.p, N * sizeof(fe::ViewIn), cudaMemcpyHostToDevice, ctx->stream));
    }
    {
        const int rc = masked_scratch();
        if (rc) return rc;
    }
    auto host_pass = [&](auto emit_tag, const fe::Out &out, bool slabs) { // the kernel's body, view by view
        constexpr bool EMIT = decltype(emit_tag)::value;
        const fe::Map m = fe_make_map(ctx, false, phases);
        for (size_t i = 0; i < N; i++) {
            fe::Bases b{0, 0, 0, 0, 0, 0, {0, 0}};
            fe::Caps cap = unlimited;
            if (EMIT && slabs) {
                const uint32_t u = (uint32_t)i;
                b = fe::Bases{u * slab.ops, u * slab.segs, u * slab.cols, u * slab.planes, u * slab.parr, (int32_t)i, {0, 0}};
                cap = slab;
            } else if (EMIT) {
                b = S.h_bases.p[i];
                if (b.frame < 0) continue;
            }
            fe::Frame<EMIT> fr(m);
            fr.sc = fe::Scratch{hs_hor.data(), hs_focl.data(), hs_cocl.data(), {hs_rows.data(), hs_rows.data() + W}, hs_order.data(),
                                reinterpret_cast<uint32_t *>(hs_order.data() + S.segs.size()), nullptr, nullptr, hs_renders.data(),
                                hs_allcols.data(), hs_dsegs.data(), hs_mos.data(), hs_mo_order.data(), hs_dseg_part.data(), cap_renders, cap_allcols, cap_dsegs, cap_mos};
            fr.out = out;
            fr.cap = cap;
            fr.run(S.h_views_in.p[i], b);
            if (!EMIT || slabs) S.h_counts.p[i] = fr.n;
        }
    };
    fe::Slabs sl{};
    sl.cap = slab;
    bool pre_done = false;
    auto pre_once = [&]() -> int { // the stateless (viewpoint, seg) kernel: once per batch, in front of the first front-end launch (and inside its timing)
        if (pre_done) return DRR_OK;
        pre_done = true;
        CU(ctx, launch_fe_pre(fe_make_map(ctx, true, phases), S.d_views_in.p, n, scr, ctx->stream));
        ctx->stats.kernel_launches++;
        return DRR_OK;
    };
    for (int attempt = 0;; attempt++) { // again when a view outgrew its slab (then two-pass) or the masked phase's working arrays (then larger ones)
        if (single) {
            if (on_host) {
                hsl_views.resize(N);
                hsl_ops.resize(N * slab.ops);
                hsl_segs.resize(N * slab.segs);
                hsl_cols.resize(N * slab.cols);
                hsl_planes.resize(N * slab.planes);
                hsl_parr.resize(N * slab.parr);
                sl.out = fe::Out{hsl_views.data(), hsl_ops.data(), hsl_segs.data(), hsl_cols.data(), hsl_planes.data(), hsl_parr.data()};
                host_pass(std::true_type{}, sl.out, true);
            } else {
                CU(ctx, S.d_sl_views.reserve(N));
                CU(ctx, S.d_sl_ops.reserve(N * slab.ops));
                CU(ctx, S.d_sl_segs.reserve(N * slab.segs));
                CU(ctx, S.d_sl_cols.reserve(N * slab.cols * 5));
                CU(ctx, S.d_sl_planes.reserve(N * slab.planes));
                CU(ctx, S.d_sl_parr.reserve(N * slab.parr));
                sl.out = fe::Out{S.d_sl_views.p, S.d_sl_ops.p, S.d_sl_segs.p, reinterpret_cast<ColRec *>(S.d_sl_cols.p), S.d_sl_planes.p, S.d_sl_parr.p};
                CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
                if (int rc = pre_once()) return rc;
                CU(ctx, launch_frontend(true, fe_make_map(ctx, true, phases), S.d_views_in.p, nullptr, S.d_counts.p, n, scr, sl.out, slab, ctx->stream));
                ctx->stats.kernel_launches++;
                CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
            }
        } else if (on_host) {
            host_pass(std::false_type{}, fe::Out{}, false);
        } else {
            CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
            if (int rc = pre_once()) return rc;
            CU(ctx, launch_frontend(false, fe_make_map(ctx, true, phases), S.d_views_in.p, nullptr, S.d_counts.p, n, scr, fe::Out{}, nocap, ctx->stream));
            ctx->stats.kernel_launches++;
            CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        }
        if (!on_host) {
            mark("front-end kernel launched");
            CU(ctx, cudaMemcpyAsync(S.h_counts.p, S.d_counts.p, N * sizeof(fe::Counts), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            mark("counts on the host");
            if (single) CU(ctx, cudaEventElapsedTime(&S.emit_ms, ctx->ev[2], ctx->ev[3]));
            else CU(ctx, cudaEventElapsedTime(&S.count_ms, ctx->ev[0], ctx->ev[1]));
        }
        bool overflow = false, scratch = false;
        for (size_t i = 0; i < N; i++) {
            overflow |= single && S.h_counts.p[i].status == fe::FE_HARD && S.h_counts.p[i].detail == fe::FED_CAPACITY;
            scratch |= S.h_counts.p[i].status == fe::FE_HARD && S.h_counts.p[i].detail == fe::FED_SCRATCH;
        }
        if (scratch && attempt < 6 && (uint64_t)N * cap_allcols * 4 * sizeof(ColRec) < (1ull << 34)) {
            cap_renders *= 4;
            cap_allcols *= 4;
            cap_dsegs *= 4;
            if (!ctx->knobs.fe_cap_renders && !ctx->knobs.fe_cap_dsegs) S.scratch_boost = std::min(64u, S.scratch_boost * 4u); // the context's next batch starts there
            const int rc = masked_scratch();
            if (rc) return rc;
            continue;
        }
        if (!overflow) break;
        single = false; // a view outgrew its slab: size the lists exactly (and give the context's next batch larger slabs)
        if (!ctx->knobs.fe_slab_div) S.slab_boost = std::min(8u, S.slab_boost * 2u);
    }
    // offsets: an exclusive scan over the viewpoints that got a frame
    uint64_t ops = 0, segs = 0, cols = 0, planes = 0, parr = 0, reccap = 0, nrec = 0;
    size_t nf = 0;
    ctx->frame_op_base.clear();
    ctx->frame_rec_base.clear();
    ctx->frame_slot.clear();
    ctx->frame_seg_base.clear();
    ctx->frame_col_base.clear();
    ctx->frame_plane_base.clear();
    ctx->frame_parr_base.clear();
    auto push_bases = [&]() {
        ctx->frame_seg_base.push_back((uint32_t)segs);
        ctx->frame_col_base.push_back((uint32_t)cols);
        ctx->frame_plane_base.push_back((uint32_t)planes);
        ctx->frame_parr_base.push_back((uint32_t)parr);
        return ctx->frame_op_base.push((uint32_t)ops) && ctx->frame_rec_base.push((uint32_t)reccap);
    };
    if (!push_bases()) return fail(ctx, DRR_E_NOMEM, "alloc");
    for (size_t i = 0; i < N; i++) {
        const fe::Counts &c = S.h_counts.p[i];
        fe::Bases &b = S.h_bases.p[i];
        b = fe::Bases{(uint32_t)ops, (uint32_t)segs, (uint32_t)cols, (uint32_t)planes, (uint32_t)parr, -1, {0, 0}};
        if (c.status == fe::FE_HARD) {
            ctx->clear_lists();
            std::fill(ctx->slot_to_frame.begin(), ctx->slot_to_frame.end(), -1);
            return fail(ctx, (c.detail == fe::FED_STACK || c.detail == fe::FED_SCRATCH) ? DRR_E_INVALID : DRR_E_ASSET,
                        std::string("drr_fe_emit_views: view ") + std::to_string(i) + ": " + fe_detail_message(c.detail));
        }
        if (status) status[i] = c.status == fe::FE_OK ? DRR_OK : DRR_E_PANIC;
        if (c.status != fe::FE_OK) continue; // the reference would have panicked: no frame for this viewpoint
        b.frame = (int32_t)nf++;
        ops += c.nops;
        segs += c.nsegs;
        cols += c.ncols;
        planes += c.nplanes;
        parr += c.nparr;
        reccap += c.reccap;
        nrec += c.nrec;
        if (!ctx->frame_slot.push((uint32_t)(first_view_idx + (int)i)) || !push_bases()) return fail(ctx, DRR_E_NOMEM, "alloc");
        ctx->slot_to_frame[first_view_idx + (int)i] = b.frame;
    }
    int nbands, band_rows;
    tile_bands(ctx->H, ctx->knobs.tile_max_rows, &nbands, &band_rows);
    const size_t nlists = nbands <= MAX_LIST_BANDS ? (size_t)nbands : 1;
    if (reccap * nlists > 0xffffffffull || cols > 0xffffffffull || parr > 0xffffffffull || ops > 0x7fffffffull) {
        ctx->clear_lists();
        std::fill(ctx->slot_to_frame.begin(), ctx->slot_to_frame.end(), -1);
        return fail(ctx, DRR_E_INVALID, "batch too large: more than 2^32 column records");
    }
    ctx->rec_cap = reccap;
    ctx->rec_count = nrec;
    ctx->stats.frames = nf;
    ctx->stats.seg_headers = segs;
    ctx->stats.column_records = cols;
    ctx->stats.visplanes = planes;
    ctx->stats.visplane_columns = parr;
    S.device_list_bytes = nf * sizeof(View) + ops * 4 + (nf + 1) * 8 + nf * 4 + segs * sizeof(SegRec) + cols * sizeof(ColRec) + planes * sizeof(PlaneRec) + parr * 4;
    S.single_pass = single;
    if (nf == 0) return DRR_OK;
    if (on_host) {
        if (!ctx->views.reserve(nf) || !ctx->ops.reserve(ops) || !ctx->segs.reserve(segs) || !ctx->cols.reserve(cols) || !ctx->planes.reserve(planes) ||
            !ctx->parr.reserve(parr))
            return fail(ctx, DRR_E_NOMEM, "alloc");
        ctx->views.n = nf;
        ctx->ops.n = ops;
        ctx->segs.n = segs;
        ctx->cols.n = cols;
        ctx->planes.n = planes;
        ctx->parr.n = parr;
        const fe::Out out{ctx->views.p, ctx->ops.p, ctx->segs.p, ctx->cols.p, ctx->planes.p, ctx->parr.p};
        if (single) {
            for (size_t i = 0; i < N; i++)
                if (S.h_bases.p[i].frame >= 0) fe::compact_view(sl, (uint32_t)i, S.h_counts.p[i], S.h_bases.p[i], out);
        } else {
            host_pass(std::true_type{}, out, false);
        }
        return DRR_OK;
    }
    mark("offsets computed");
    S.slab_draw = false;
    S.dense_valid = true;
    if (single && !ctx->knobs.fe_compact) {
        // No compaction: the draw kernels read the slabs (the ops, SegRec::cols_first and PlaneRec::arr_first in them are indices into
        // the slab arrays already).  What they need per frame: where its ops start and how many there are, its record slots, its
        // framebuffer slot, and its View in an array indexed by frame.
        if (!S.h_slab_meta.reserve(2 * nf + 1)) return fail(ctx, DRR_E_NOMEM, "alloc");
        size_t f = 0;
        for (size_t i = 0; i < N; i++)
            if (S.h_bases.p[i].frame >= 0) {
                S.h_slab_meta.p[f] = (uint32_t)i * slab.ops;
                S.h_slab_meta.p[nf + 1 + f] = S.h_counts.p[i].nops;
                f++;
            }
        S.h_slab_meta.p[nf] = 0u;
        CU(ctx, ctx->d_views.reserve(nf));
        CU(ctx, S.d_slab_meta.reserve(2 * nf + 1));
        CU(ctx, ctx->d_frame_rec_base.reserve(nf + 1));
        CU(ctx, ctx->d_frame_slot.reserve(nf));
        CU(ctx, ctx->d_frame_cursor.reserve(nf));
        CU(ctx, ctx->d_colidx.reserve(nf * W * nlists));
        CU(ctx, ctx->d_tparams.reserve(std::max<uint64_t>(reccap, 1) * 4 * nlists));
        CU(ctx, cudaMemcpyAsync(S.d_slab_meta.p, S.h_slab_meta.p, (2 * nf + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->d_frame_rec_base.p, ctx->frame_rec_base.p, (nf + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->d_frame_slot.p, ctx->frame_slot.p, nf * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        CU(ctx, launch_fe_gather_views(sl.out.views, ctx->d_frame_slot.p, first_view_idx, (int)nf, ctx->d_views.p, ctx->stream));
        ctx->stats.kernel_launches++;
        CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        S.slab_draw = true;
        S.dense_valid = false;
        S.sl_last = sl;
        S.n_last = n;
        S.first_last = first_view_idx;
        ctx->device_lists = true;
        ctx->uploaded_frames = nf;
        mark("views gathered, slabs handed to the draw kernels");
        return DRR_OK;
    }
    // the dense device lists: sized from the counts
    CU(ctx, ctx->d_views.reserve(nf));
    CU(ctx, ctx->d_ops.reserve(std::max<uint64_t>(ops, 1)));
    CU(ctx, ctx->d_frame_op_base.reserve(nf + 1));
    CU(ctx, ctx->d_frame_rec_base.reserve(nf + 1));
    CU(ctx, ctx->d_frame_slot.reserve(nf));
    CU(ctx, ctx->d_frame_cursor.reserve(nf));
    CU(ctx, ctx->d_segs.reserve(std::max<uint64_t>(segs, 1)));
    CU(ctx, ctx->d_cols.reserve(std::max<uint64_t>(cols, 1) * 5));
    CU(ctx, ctx->d_planes.reserve(std::max<uint64_t>(planes, 1)));
    CU(ctx, ctx->d_parr.reserve(std::max<uint64_t>(parr, 1)));
    CU(ctx, ctx->d_colidx.reserve(nf * W * nlists));
    CU(ctx, ctx->d_tparams.reserve(std::max<uint64_t>(reccap, 1) * 4 * nlists));
    CU(ctx, cudaMemcpyAsync(S.d_bases.p, S.h_bases.p, N * sizeof(fe::Bases), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_op_base.p, ctx->frame_op_base.p, (nf + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_rec_base.p, ctx->frame_rec_base.p, (nf + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_frame_slot.p, ctx->frame_slot.p, nf * 4, cudaMemcpyHostToDevice, ctx->stream));
    const fe::Out out{ctx->d_views.p, ctx->d_ops.p, ctx->d_segs.p, reinterpret_cast<ColRec *>(ctx->d_cols.p), ctx->d_planes.p, ctx->d_parr.p};
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if (single) CU(ctx, launch_fe_compact(sl, S.d_counts.p, S.d_bases.p, n, out, ctx->stream));
    else CU(ctx, launch_frontend(true, fe_make_map(ctx, true, phases), S.d_views_in.p, S.d_bases.p, nullptr, n, scr, out, nocap, ctx->stream));
    ctx->stats.kernel_launches++;
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->device_lists = true;
    ctx->uploaded_frames = nf;
    mark("compaction / emit pass launched");
    return DRR_OK;
}

int drr_fe_emit_views(drr_ctx *ctx, int first_view_idx, const float *xya, int n, int phases, int *status) {
    CTX_DEV(ctx);
    return fe_emit_views(ctx, first_view_idx, xya, n, phases, status, false);
}

int drr_fe_last_times(drr_ctx *ctx, float *count_ms, float *emit_ms) {
    CTX_DEV(ctx);
    if (ctx->host_only || !ctx->device_lists) return fail(ctx, DRR_E_STATE, "drr_fe_last_times: no device front-end batch");
    FeState &S = ctx->fes;
    CU(ctx, cudaEventSynchronize(ctx->ev[1]));
    float t = 0;
    CU(ctx, cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1])); // the batch's last kernel: compaction (single-pass) or the emit pass (two-pass)
    if (S.single_pass) {
        if (count_ms) *count_ms = t;          // single-pass mode has no count pass: the second figure's slot carries the compaction
        if (emit_ms) *emit_ms = S.emit_ms;
    } else {
        if (count_ms) *count_ms = S.count_ms;
        if (emit_ms) *emit_ms = t;
    }
    return DRR_OK;
}
int drr_fe_last_mode(drr_ctx *ctx) { return ctx ? (ctx->fes.single_pass ? 1 : 2) : DRR_E_INVALID; }

// ==== test infrastructure: compiled only into libdrr_test.so (-DDRR_TESTING), never into the product library ===============
#ifdef DRR_TESTING
// Test infrastructure: the per-view front-end code of drr_frontend.cuh run on the CPU into the context's host lists
// (works in a recording-only context), so that tests/ can compare it list by list with the host front-end.
int drr_test_fe_emit_views_host(drr_ctx *ctx, int first_view_idx, const float *xya, int n, int phases, int *status) {
    CTX_CHECK(ctx);
    return fe_emit_views(ctx, first_view_idx, xya, n, phases, status, true);
}
// Test infrastructure: copy the lists the device front-end wrote back into the context's host lists (drr_test_list).
int drr_test_fe_download_lists(drr_ctx *ctx) {
    CTX_DEV(ctx);
    if (ctx->host_only || !ctx->device_lists) return fail(ctx, DRR_E_STATE, "drr_test_fe_download_lists: no device front-end batch");
    const size_t nf = ctx->uploaded_frames;
    const size_t ops = ctx->frame_op_base.p[nf], segs = ctx->frame_seg_base[nf], cols = ctx->frame_col_base[nf], planes = ctx->frame_plane_base[nf],
                 parr = ctx->frame_parr_base[nf];
    if (!ctx->views.reserve(nf) || !ctx->ops.reserve(ops) || !ctx->segs.reserve(segs) || !ctx->cols.reserve(cols) || !ctx->planes.reserve(planes) ||
        !ctx->parr.reserve(parr))
        return fail(ctx, DRR_E_NOMEM, "alloc");
    FeState &S = ctx->fes;
    if (S.slab_draw && !S.dense_valid) { // the batch was drawn from its slabs: make the dense lists now (what DRR_FE_COMPACT=1 does in line)
        CU(ctx, ctx->d_ops.reserve(std::max<size_t>(ops, 1)));
        CU(ctx, ctx->d_segs.reserve(std::max<size_t>(segs, 1)));
        CU(ctx, ctx->d_cols.reserve(std::max<size_t>(cols, 1) * 5));
        CU(ctx, ctx->d_planes.reserve(std::max<size_t>(planes, 1)));
        CU(ctx, ctx->d_parr.reserve(std::max<size_t>(parr, 1)));
        CU(ctx, cudaMemcpyAsync(S.d_bases.p, S.h_bases.p, (size_t)S.n_last * sizeof(fe::Bases), cudaMemcpyHostToDevice, ctx->stream));
        const fe::Out out{ctx->d_views.p, ctx->d_ops.p, ctx->d_segs.p, reinterpret_cast<ColRec *>(ctx->d_cols.p), ctx->d_planes.p, ctx->d_parr.p};
        CU(ctx, launch_fe_compact(S.sl_last, S.d_counts.p, S.d_bases.p, S.n_last, out, ctx->stream));
        S.dense_valid = true;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaMemcpy(ctx->views.p, ctx->d_views.p, nf * sizeof(View), cudaMemcpyDeviceToHost));
    if (ops) CU(ctx, cudaMemcpy(ctx->ops.p, ctx->d_ops.p, ops * 4, cudaMemcpyDeviceToHost));
    if (segs) CU(ctx, cudaMemcpy(ctx->segs.p, ctx->d_segs.p, segs * sizeof(SegRec), cudaMemcpyDeviceToHost));
    if (cols) CU(ctx, cudaMemcpy(ctx->cols.p, ctx->d_cols.p, cols * sizeof(ColRec), cudaMemcpyDeviceToHost));
    if (planes) CU(ctx, cudaMemcpy(ctx->planes.p, ctx->d_planes.p, planes * sizeof(PlaneRec), cudaMemcpyDeviceToHost));
    if (parr) CU(ctx, cudaMemcpy(ctx->parr.p, ctx->d_parr.p, parr * 4, cudaMemcpyDeviceToHost));
    ctx->views.n = nf;
    ctx->ops.n = ops;
    ctx->segs.n = segs;
    ctx->cols.n = cols;
    ctx->planes.n = planes;
    ctx->parr.n = parr;
    return DRR_OK;
}

// ---- CPU-testable internals (no CUDA needed) ---------------------------------------------------------------------------
// A context that can RECORD draw lists without a GPU (plain malloc staging) so the host logic is testable in the CPU-only
// container.  It has no framebuffers and every execution entry point fails with DRR_E_CUDA.
int drr_test_ctx_create_host_only(int width, int height, int max_views, drr_ctx **out) {
    if (!out || width <= 0 || height <= 0 || width > 16384 || height > 16384 || max_views <= 0) return DRR_E_INVALID;
    drr_ctx *c = new drr_ctx();
    c->knobs.read_environment();
    c->host_only = true;
    c->W = width;
    c->H = height;
    c->max_views = max_views;
    c->ASPECT = 200.0f / 240.0f;
    c->GCFX = ((float)(uint32_t)width / c->ASPECT) / 2.0f;
    c->CFX = (float)(uint32_t)width / 2.0f;
    c->CFY = (float)(uint32_t)height / 2.0f;
    c->slot_to_frame.assign(max_views, -1);
    c->set_pinned(false);
    c->h_crc.pinned = false;
    c->fes.h_views_in.pinned = c->fes.h_counts.pinned = c->fes.h_bases.pinned = false;
    *out = c;
    return DRR_OK;
}

// Host restatement of the bin kernel's column binning (drr_tile.cu: walk_column), op by op in call order: the per-column
// span lists the device is expected to build.  Test infrastructure only (tests/ compare the device's lists with these,
// and replay them on the CPU with the reference restatement's leaf drawers); the product never draws from them.
static void host_reference_binning(drr_ctx *ctx) {
    const size_t nf = ctx->frame_op_base.n ? ctx->frame_op_base.n - 1 : 0;
    const int W = ctx->W, H = ctx->H;
    ctx->t_spans.clear();
    ctx->t_colidx.assign(nf * (size_t)W, ColIdx{0, 0});
    std::vector<std::vector<Span>> percol(W);
    for (size_t f = 0; f < nf; f++) {
        for (auto &v : percol) v.clear();
        for (uint32_t o = ctx->frame_op_base.p[f]; o < ctx->frame_op_base.p[f + 1]; o++) {
            const uint32_t op = ctx->ops.p[o];
            if (op & 0x80000000u) {
                const PlaneRec &p = ctx->planes.p[op & 0x7fffffffu];
                for (int x = p.left; x <= p.right; x++) {
                    const uint32_t tb = ctx->parr.p[p.arr_first + (uint32_t)(x - p.left)];
                    const int t = std::max<int>((int16_t)(tb & 0xffffu), 0), b = std::min<int>((int16_t)(tb >> 16), H - 1);
                    if (p.kind == (int16_t)KIND_FLAT && (int16_t)(b - t) <= 1) continue;
                    if (t > b) continue;
                    percol[x].push_back(Span{(uint16_t)t, (uint16_t)b, (uint16_t)x, (uint8_t)p.kind, 0, op & 0x7fffffffu, 0, 0});
                }
            } else {
                const SegRec &g = ctx->segs.p[op];
                const uint8_t kind = ctx->bitmaps[g.bitmap_slot].opaque ? KIND_WALL : KIND_WALL_HOLES;
                for (uint32_t i = 0; i < g.n; i++) {
                    const ColRec &c = ctx->cols.p[g.cols_first + i];
                    if (c.x < 0 || c.x >= W) continue;
                    const int a = std::max<int>(c.clipped_top_y, 0), b = std::min<int>(c.clipped_bottom_y, H - 1);
                    if (a > b) continue;
                    percol[c.x].push_back(Span{(uint16_t)a, (uint16_t)b, (uint16_t)c.x, kind, 0, op, c.top_y, c.bottom_y});
                }
            }
        }
        for (int x = 0; x < W; x++) {
            ctx->t_colidx[f * (size_t)W + x] = ColIdx{(uint32_t)ctx->t_spans.size(), (uint32_t)percol[x].size()};
            ctx->t_spans.insert(ctx->t_spans.end(), percol[x].begin(), percol[x].end());
        }
    }
}

// Raw views of the recorded lists: which = 0 views, 1 segs, 2 planes, 3 spans and 4 colidx of the host reference binning,
// 5 frame_rec_base, 6 frame_slot, 7 column records, 8 visplane (top, bottom) pairs, 9 ops, 10 frame_op_base
const void *drr_test_list(drr_ctx *ctx, int which, uint64_t *count, uint64_t *elem_size) {
    if (!ctx || !count || !elem_size) return nullptr;
    if ((which == 3 || which == 4) && ctx->t_colidx.empty() && !ctx->in_frame) host_reference_binning(ctx);
    switch (which) {
    case 0: *count = ctx->views.n; *elem_size = sizeof(View); return ctx->views.p;
    case 1: *count = ctx->segs.n; *elem_size = sizeof(SegRec); return ctx->segs.p;
    case 2: *count = ctx->planes.n; *elem_size = sizeof(PlaneRec); return ctx->planes.p;
    case 3: *count = ctx->t_spans.size(); *elem_size = sizeof(Span); return ctx->t_spans.data();
    case 4: *count = ctx->t_colidx.size(); *elem_size = sizeof(ColIdx); return ctx->t_colidx.data();
    case 5: *count = ctx->frame_rec_base.n; *elem_size = 4; return ctx->frame_rec_base.p;
    case 6: *count = ctx->frame_slot.n; *elem_size = 4; return ctx->frame_slot.p;
    case 7: *count = ctx->cols.n; *elem_size = sizeof(ColRec); return ctx->cols.p;
    case 8: *count = ctx->parr.n; *elem_size = 4; return ctx->parr.p;
    case 9: *count = ctx->ops.n; *elem_size = 4; return ctx->ops.p;
    case 10: *count = ctx->frame_op_base.n; *elem_size = 4; return ctx->frame_op_base.p;
    default: return nullptr;
    }
}
// What the bin kernel produced for the uploaded batch: colidx_out = nframes * nlists * W (first, n) pairs (one list per row
// band, drr_test_tile_bands), recs_out = two words per record slot (y0 | y1 << 16, kind | flags), rec_cap * nlists slots,
// indexed by the `first` values of colidx_out.
int drr_test_device_bins(drr_ctx *ctx, uint32_t *colidx_out, uint32_t *recs_out) {
    CTX_DEV(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    if (!colidx_out || !recs_out || ctx->uploaded_frames == 0) return fail(ctx, DRR_E_STATE, "drr_test_device_bins: nothing drawn");
    int nbands, band_rows;
    tile_bands(ctx->H, ctx->knobs.tile_max_rows, &nbands, &band_rows);
    const size_t nlists = nbands <= MAX_LIST_BANDS ? (size_t)nbands : 1;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaMemcpy(colidx_out, ctx->d_colidx.p, ctx->uploaded_frames * nlists * (size_t)ctx->W * sizeof(ColIdx), cudaMemcpyDeviceToHost));
    if (ctx->rec_cap)
        CU(ctx, cudaMemcpy2D(recs_out, 8, ctx->d_tparams.p, 64, 8, (size_t)ctx->rec_cap * nlists, cudaMemcpyDeviceToHost));
    return DRR_OK;
}
// how the tile kernel cuts this context's columns into row bands, and how many span lists per column the bin kernel writes
int drr_test_tile_bands(drr_ctx *ctx, int *nbands, int *band_rows, int *nlists) {
    if (!ctx || !nbands || !band_rows || !nlists) return DRR_E_INVALID;
    tile_bands(ctx->H, ctx->knobs.tile_max_rows, nbands, band_rows);
    *nlists = *nbands <= MAX_LIST_BANDS ? *nbands : 1;
    return DRR_OK;
}
// bitmap slot -> (w, h, opaque); flat_slot/bitmap_slot resolve ids the way the device tables do
int drr_test_bitmap_info(drr_ctx *ctx, int slot, int *w, int *h, int *opaque) {
    if (!ctx || slot < 0 || slot >= (int)ctx->bitmaps.size()) return DRR_E_INVALID;
    *w = ctx->bitmaps[slot].w;
    *h = ctx->bitmaps[slot].h;
    *opaque = (int)ctx->bitmaps[slot].opaque;
    return DRR_OK;
}
int drr_test_bitmap_texels(drr_ctx *ctx, int slot, int16_t *out) { // row-major w*h, -1 = None (decoded back from the device pool layout)
    if (!ctx || slot < 0 || slot >= (int)ctx->bitmaps.size()) return DRR_E_INVALID;
    const BitmapRec &r = ctx->bitmaps[slot];
    uint32_t pitch = 1;
    while (pitch < (uint32_t)r.h) pitch <<= 1;
    for (int y = 0; y < r.h; y++)
        for (int x = 0; x < r.w; x++) {
            const uint16_t t = ctx->texel_pool[r.base + (size_t)x * pitch + y];
            const uint32_t idx = ((uint32_t)t - (ctx->pal_base + SM_PAL)) / PAL_ENTRY;
            out[(size_t)y * r.w + x] = idx == TEXEL_NONE_INDEX ? (int16_t)-1 : (int16_t)idx;
        }
    return DRR_OK;
}
int drr_test_flat_texels(drr_ctx *ctx, int slot, uint8_t *out4096) {
    if (!ctx || slot < 0 || (size_t)(slot + 1) * 4096 > ctx->flat_pool.size()) return DRR_E_INVALID;
    memcpy(out4096, ctx->flat_pool.data() + (size_t)slot * 4096, 4096);
    return DRR_OK;
}
int drr_test_palette(drr_ctx *ctx, uint8_t *out768) {
    if (!ctx || !ctx->have_pal) return DRR_E_INVALID;
    for (int i = 0; i < 256; i++) {
        out768[i * 3] = (uint8_t)ctx->pal[i].x;
        out768[i * 3 + 1] = (uint8_t)ctx->pal[i].y;
        out768[i * 3 + 2] = (uint8_t)ctx->pal[i].z;
    }
    return DRR_OK;
}
int drr_test_sky_slot(drr_ctx *ctx) { return ctx ? ctx->sky_slot : -1; }
int drr_test_tile_config(drr_ctx *ctx, int *tc, int *lpg) {
    if (!ctx || !tc || !lpg) return DRR_E_INVALID;
    tile_config(ctx->W, ctx->H, tc, lpg);
    return DRR_OK;
}
int drr_test_bitmap_id_of_slot(drr_ctx *ctx, int slot) {
    for (auto &kv : ctx->bitmap_slot)
        if (kv.second == slot) return kv.first;
    return -1;
}
int drr_test_flat_id_of_slot(drr_ctx *ctx, int slot) {
    for (auto &kv : ctx->flat_slot)
        if (kv.second == slot) return kv.first;
    return -1;
}

// Device self-check of the hoisted-reciprocal division used by the pixel loops (drr_math.cuh: fast_div) against
// __fdiv_rn.  mode 0: a in [-amax, amax] (n0 = 2*amax+1 values), b = all non-zero integers in [-n1/2, n1/2];
// mode 1: a = floats with bit patterns lo + k*stride (k < n0), b = CFY - y for y < n1 (H = n1).  Returns mismatches.
int drr_test_fastdiv(drr_ctx *ctx, int mode, long long n0, long long n1, float CFY, uint32_t lo, uint32_t stride, unsigned long long *bad,
                     float *first2) {
    CTX_CHECK(ctx);
    if (ctx->host_only) return fail(ctx, DRR_E_CUDA, "recording-only test context");
    unsigned long long *d_bad = nullptr;
    float *d_first = nullptr;
    CU(ctx, cudaMalloc((void **)&d_bad, 8));
    CU(ctx, cudaMalloc((void **)&d_first, 8));
    CU(ctx, cudaMemsetAsync(d_bad, 0, 8, ctx->stream));
    CU(ctx, cudaMemsetAsync(d_first, 0, 8, ctx->stream));
    CU(ctx, launch_fastdiv_check(mode, n0, n1, CFY, (int)n1, lo, stride, d_bad, d_first, ctx->stream));
    CU(ctx, cudaMemcpyAsync(bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(first2, d_first, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bad);
    cudaFree(d_first);
    return DRR_OK;
}

#endif // DRR_TESTING

} // extern "C"
