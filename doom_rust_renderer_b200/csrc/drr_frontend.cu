// drr_frontend.cu -- drr_frontend_kernel: the reference's front-end on the device, one WARP per viewpoint (see
// drr_frontend.cuh for what it restates and why the same source is also compiled for the CPU test harness).
//
// A viewpoint's walk is sequential from seg to seg (the occlusion arrays and the open visplanes carry state in BSP order),
// but viewpoints are independent and so are the screen columns inside a seg and the visibility tests of a subsector's
// segs: the batch is the outer parallel axis (a few thousand warps over all SMs), the lanes take columns / segs.  The
// per-view state (three W-entry occlusion arrays, two W-entry visplane row buffers) lives in global scratch
// and stays in L1/L2.  The kernel runs twice per batch: COUNT sizes every view's lists, the host turns the counts into
// offsets (exclusive scan over a few thousand entries), EMIT writes ops / SegRec / ColRec / PlaneRec / (top, bottom) pairs
// straight into the arrays drr_bin_kernel reads -- no draw list ever crosses PCIe.
#include "drr_frontend.cuh"
#include "drr_kernels.h"
#include <algorithm>
#include <new>

namespace drr {

// slab.ops != 0: single-pass mode -- no count pass ran, view v writes into its own slab of every array (drr_frontend.cuh:
// Slabs) and leaves its counts; drr_fe_compact_kernel then makes the lists dense.
template <bool EMIT, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) drr_frontend_kernel(const __grid_constant__ fe::Map m, const fe::ViewIn *__restrict__ views, const fe::Bases *__restrict__ bases,
                                                                                 fe::Counts *__restrict__ counts, int n, FeScratch s, fe::Out out, fe::Caps slab,
                                                                                 int smem_mode, uint32_t smem_per_view) {
    extern __shared__ __align__(16) uint8_t fe_smem[];
    constexpr size_t frame_bytes = (sizeof(fe::Frame<EMIT>) + 15) & ~(size_t)15;
    const int v = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); // one warp per viewpoint
    if (v >= n) return;
    fe::Bases b = fe::Bases{0, 0, 0, 0, 0, 0, {0, 0}};
    fe::Caps cap = fe::Caps{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (EMIT) {
        if (slab.ops) {
            const uint32_t u = (uint32_t)v;
            b = fe::Bases{u * slab.ops, u * slab.segs, u * slab.cols, u * slab.planes, u * slab.parr, v, {0, 0}};
            cap = slab;
        } else {
            b = bases[v];
            if (b.frame < 0) return; // the reference panics on this viewpoint: no frame
        }
    }
    // The per-view state (fe::Frame: cursors, counts, pointers, open visplanes, the occlusion bit masks) is ONE object per warp in
    // shared memory.  As a local variable it is replicated per lane (it is handed to non-inlined member functions, so it lives
    // in local memory): 32 KB per warp, 900 KB per SM with every warp resident -- a third of the loads of uniform state then
    // missed L1.  Every lane executes the same uniform updates on the same shared word.
    fe::Frame<EMIT> &fr = *new (fe_smem + (size_t)(threadIdx.x >> 5) * smem_per_view) fe::Frame<EMIT>(m);
    const size_t o = (size_t)v * (size_t)m.W;
    fr.sc.hor_ocl = s.hor_ocl + o;
    fr.sc.floor_ocl = s.floor_ocl + o;
    fr.sc.ceil_ocl = s.ceil_ocl + o;
    fr.sc.rows[0] = s.rows + 2 * o;
    fr.sc.rows[1] = s.rows + 2 * o + m.W;
    fr.sc.order = s.order + (size_t)v * (size_t)(m.nsegs + m.side_words);
    fr.sc.side = reinterpret_cast<uint32_t *>(fr.sc.order + m.nsegs);
    // The per-view state the column loops hammer on goes to shared memory when it fits with every warp of the SM resident
    // (launch_frontend picks the mode): 1 = the three occlusion arrays and the side bits, 2 = also the two visplane row buffers.
    if ((smem_mode & 3) >= 1) {
        uint8_t *base = fe_smem + (size_t)(threadIdx.x >> 5) * smem_per_view + frame_bytes;
        const size_t Wp = ((size_t)m.W + 15) & ~(size_t)15;
        fr.sc.hor_ocl = base;
        fr.sc.floor_ocl = reinterpret_cast<int16_t *>(base + Wp);
        fr.sc.ceil_ocl = reinterpret_cast<int16_t *>(base + 3 * Wp);
        fr.sc.side = reinterpret_cast<uint32_t *>(base + 5 * Wp);
        uint8_t *end = base + 5 * Wp + 4 * (size_t)((m.side_words + 3) & ~3);
        if ((smem_mode & 3) >= 2) {
            fr.sc.rows[0] = reinterpret_cast<uint32_t *>(end);
            fr.sc.rows[1] = fr.sc.rows[0] + Wp;
            end += 8 * Wp;
        }
        if (smem_mode & 4) fr.sc.order = reinterpret_cast<int32_t *>(end); // the seg order too (written once, read once, by this warp)
    }
    fr.sc.pre = s.pre ? static_cast<const fe::SegPre *>(s.pre) + (size_t)v * (size_t)m.nsegs : nullptr;
    fr.sc.pre_code = s.pre_code ? s.pre_code + (size_t)v * (size_t)m.nsegs : nullptr;
    fr.sc.renders = static_cast<fe::RenderRec *>(s.renders) + (size_t)v * s.cap_renders;
    fr.sc.allcols = static_cast<ColRec *>(s.allcols) + (size_t)v * s.cap_allcols;
    fr.sc.dsegs = static_cast<SegRec *>(s.dsegs) + (size_t)v * s.cap_dsegs;
    fr.sc.mos = static_cast<fe::MoRec *>(s.mos) + (size_t)v * s.cap_mos;
    fr.sc.mo_order = s.mo_order + (size_t)v * s.cap_mos;
    fr.sc.dseg_part = s.dseg_part + (size_t)v * s.cap_dsegs;
    fr.sc.cap_renders = s.cap_renders;
    fr.sc.cap_allcols = s.cap_allcols;
    fr.sc.cap_dsegs = s.cap_dsegs;
    fr.sc.cap_mos = s.cap_mos;
    fr.out = out;
    fr.cap = cap;
    fr.run(views[v], b);
    if ((!EMIT || slab.ops) && (threadIdx.x & 31u) == 0u) counts[v] = fr.n;
}

// One thread per (viewpoint, seg): the part of process_seg that needs neither order nor state.  Consecutive threads take
// consecutive segs of one viewpoint, so the records and codes go out coalesced; ~85 % of the pairs end at the field-of-view clip.
__global__ void __launch_bounds__(256) drr_fe_pre_kernel(const __grid_constant__ fe::Map m, const fe::ViewIn *__restrict__ views, int n, fe::SegPre *__restrict__ pre,
                                                         uint8_t *__restrict__ code) {
    const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (s >= m.nsegs) return;
    for (int v = (int)blockIdx.y; v < n; v += (int)gridDim.y) {
        const fe::ViewIn vi = views[v];
        const fe::SegPre p = fe::seg_pre_of(m, fe::V2{vi.x, vi.y}, vi.cos_n, vi.sin_n, m.segs[s]);
        const size_t at = (size_t)v * (size_t)m.nsegs + (size_t)s;
        code[at] = (uint8_t)p.code;
        if (p.code) pre[at] = p;
    }
}

__global__ void __launch_bounds__(128) drr_fe_compact_kernel(fe::Slabs sl, const fe::Counts *__restrict__ counts, const fe::Bases *__restrict__ bases, int n, fe::Out dst) {
    const int v = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); // one warp per viewpoint
    if (v >= n) return;
    const fe::Bases b = bases[v];
    if (b.frame < 0) return;
    fe::compact_view(sl, (uint32_t)v, counts[v], b, dst);
}

__global__ void __launch_bounds__(256) drr_fe_gather_views_kernel(const View *__restrict__ slab_views, const uint32_t *__restrict__ frame_slot, int first_view_idx, int nframes,
                                                                  View *__restrict__ dst) {
    const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (f < nframes) dst[f] = slab_views[(int)frame_slot[f] - first_view_idx];
}

cudaError_t launch_frontend(bool emit, const fe::Map &m, const fe::ViewIn *views, const fe::Bases *bases, fe::Counts *counts, int n,
                            const FeScratch &s, const fe::Out &out, const fe::Caps &slab, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    // shared memory per viewpoint: the per-view state itself (always), + occlusion arrays 5 bytes per column + side bits (mode 1),
    // + 8 bytes per column of visplane rows (mode 2); taken when every warp the register budget allows on an SM fits in ~200 KB
    const size_t Wp = ((size_t)m.W + 15) & ~(size_t)15, side_bytes = 4 * (size_t)((m.side_words + 3) & ~3);
    const size_t frame_bytes = (std::max(sizeof(fe::Frame<true>), sizeof(fe::Frame<false>)) + 15) & ~(size_t)15;
    const size_t need1 = frame_bytes + 5 * Wp + side_bytes, need2 = need1 + 8 * Wp;
    // Eight warps per CTA when map objects are drawn (eight neighbouring viewpoints see the same sprites and masked parts: things640
    // 1.56 -> 1.39 ms per 4096 viewpoints; without map objects four warps are 1-5 % faster) and their occlusion arrays fit.
#ifdef DRR_FE_EIGHT_ALWAYS // A/B: whenever they fit
    const bool eight = need1 <= 48 * 1024 / FE_WARPS_THINGS;
#else
    const bool eight = (m.phases & 4) && m.nthings > 0 && need1 <= 48 * 1024 / FE_WARPS_THINGS;
#endif
    const int warps = eight ? FE_WARPS_THINGS : FE_WARPS, minb = eight ? FE_MIN_BLOCKS_THINGS : FE_MIN_BLOCKS;
    // (at most 48 KB per CTA: the default limit of dynamic shared memory, no opt-in needed)
    const size_t budget = std::min<size_t>(200 * 1024 / (size_t)(minb * warps), 48 * 1024 / (size_t)warps);
    // (mode 1 only up to 8 KB per viewpoint: at 1920 columns 16 viewpoints' arrays would take 170 KB of the SM and leave the
    // global scratch of the masked phase no L1 to speak of -- measured on the stress map: 64.0 ms against 56.1 ms in mode 0)
#ifdef DRR_FE_MAX_MODE // A/B: never more than this mode
    const int mode = std::min(DRR_FE_MAX_MODE, need2 <= budget ? 2 : need1 <= std::min(budget, (size_t)8192) ? 1 : 0);
#else
    const int mode = need2 <= budget ? 2 : need1 <= std::min(budget, (size_t)8192) ? 1 : 0;
#endif
    size_t per_view = mode == 2 ? need2 : mode == 1 ? need1 : frame_bytes;
    const size_t order_bytes = (4 * (size_t)m.nsegs + 15) & ~(size_t)15;
    const bool order_too = mode >= 1 && per_view + order_bytes <= budget;
    if (order_too) per_view += order_bytes;
    // mode 0 (the arrays in global scratch): the build with the smaller register budget and twice the resident warps
    const int vpb = mode == 0 ? FE_WARPS_GLOBAL : warps; // viewpoints per CTA
    const unsigned blocks = (unsigned)((n + vpb - 1) / vpb);
    const size_t dyn = per_view * vpb;
    const int mflags = mode | (order_too ? 4 : 0);
#define DRR_FE_LAUNCH(E, WARPS, MINB) drr_frontend_kernel<E, WARPS, MINB><<<blocks, 32 * vpb, dyn, st>>>(m, views, bases, counts, n, s, out, slab, mflags, (uint32_t)per_view)
    if (mode == 0) {
        if (emit) DRR_FE_LAUNCH(true, FE_WARPS_GLOBAL, FE_MIN_BLOCKS_GLOBAL);
        else DRR_FE_LAUNCH(false, FE_WARPS_GLOBAL, FE_MIN_BLOCKS_GLOBAL);
    } else if (eight) {
        if (emit) DRR_FE_LAUNCH(true, FE_WARPS_THINGS, FE_MIN_BLOCKS_THINGS);
        else DRR_FE_LAUNCH(false, FE_WARPS_THINGS, FE_MIN_BLOCKS_THINGS);
    } else {
        if (emit) DRR_FE_LAUNCH(true, FE_WARPS, FE_MIN_BLOCKS);
        else DRR_FE_LAUNCH(false, FE_WARPS, FE_MIN_BLOCKS);
    }
#undef DRR_FE_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_fe_pre(const fe::Map &m, const fe::ViewIn *views, int n, const FeScratch &s, cudaStream_t st) {
    if (n <= 0 || !s.pre || !s.pre_code) return cudaSuccess;
    const dim3 grid((unsigned)((m.nsegs + 255) / 256), (unsigned)std::min(n, 65535));
    drr_fe_pre_kernel<<<grid, 256, 0, st>>>(m, views, n, static_cast<fe::SegPre *>(s.pre), s.pre_code);
    return cudaGetLastError();
}

cudaError_t launch_fe_gather_views(const View *slab_views, const uint32_t *frame_slot, int first_view_idx, int nframes, View *dst, cudaStream_t st) {
    if (nframes <= 0) return cudaSuccess;
    drr_fe_gather_views_kernel<<<(unsigned)((nframes + 255) / 256), 256, 0, st>>>(slab_views, frame_slot, first_view_idx, nframes, dst);
    return cudaGetLastError();
}

cudaError_t launch_fe_compact(const fe::Slabs &sl, const fe::Counts *counts, const fe::Bases *bases, int n, const fe::Out &dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    drr_fe_compact_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(sl, counts, bases, n, dst);
    return cudaGetLastError();
}

} // namespace drr
