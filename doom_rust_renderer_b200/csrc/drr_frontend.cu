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

namespace drr {

// slab.ops != 0: single-pass mode -- no count pass ran, view v writes into its own slab of every array (drr_frontend.cuh:
// Slabs) and leaves its counts; drr_fe_compact_kernel then makes the lists dense.
template <bool EMIT>
__global__ void __launch_bounds__(FE_THREADS, FE_MIN_BLOCKS) drr_frontend_kernel(const __grid_constant__ fe::Map m, const fe::ViewIn *__restrict__ views, const fe::Bases *__restrict__ bases,
                                                                                 fe::Counts *__restrict__ counts, int n, FeScratch s, fe::Out out, fe::Caps slab) {
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
    fe::Frame<EMIT> fr(m);
    const size_t o = (size_t)v * (size_t)m.W;
    fr.sc.hor_ocl = s.hor_ocl + o;
    fr.sc.floor_ocl = s.floor_ocl + o;
    fr.sc.ceil_ocl = s.ceil_ocl + o;
    fr.sc.rows[0] = s.rows + 2 * o;
    fr.sc.rows[1] = s.rows + 2 * o + m.W;
    fr.sc.order = s.order + (size_t)v * (size_t)m.nsegs;
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

__global__ void __launch_bounds__(128) drr_fe_compact_kernel(fe::Slabs sl, const fe::Counts *__restrict__ counts, const fe::Bases *__restrict__ bases, int n, fe::Out dst) {
    const int v = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); // one warp per viewpoint
    if (v >= n) return;
    const fe::Bases b = bases[v];
    if (b.frame < 0) return;
    fe::compact_view(sl, (uint32_t)v, counts[v], b, dst);
}

cudaError_t launch_frontend(bool emit, const fe::Map &m, const fe::ViewIn *views, const fe::Bases *bases, fe::Counts *counts, int n,
                            const FeScratch &s, const fe::Out &out, const fe::Caps &slab, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int vpb = FE_THREADS / 32; // viewpoints per CTA
    const unsigned blocks = (unsigned)((n + vpb - 1) / vpb);
    if (emit)
        drr_frontend_kernel<true><<<blocks, FE_THREADS, 0, st>>>(m, views, bases, counts, n, s, out, slab);
    else
        drr_frontend_kernel<false><<<blocks, FE_THREADS, 0, st>>>(m, views, bases, counts, n, s, out, slab);
    return cudaGetLastError();
}

cudaError_t launch_fe_compact(const fe::Slabs &sl, const fe::Counts *counts, const fe::Bases *bases, int n, const fe::Out &dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    drr_fe_compact_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(sl, counts, bases, n, dst);
    return cudaGetLastError();
}

} // namespace drr
