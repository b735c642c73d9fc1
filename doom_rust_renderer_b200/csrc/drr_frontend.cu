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

template <bool EMIT>
__global__ void __launch_bounds__(FE_THREADS) drr_frontend_kernel(fe::Map m, const fe::ViewIn *__restrict__ views, const fe::Bases *__restrict__ bases,
                                                                  fe::Counts *__restrict__ counts, int n, FeScratch s, fe::Out out) {
    const int v = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); // one warp per viewpoint
    if (v >= n) return;
    fe::Bases b = fe::Bases{0, 0, 0, 0, 0, 0, 0, 0};
    if (EMIT) {
        b = bases[v];
        if (b.frame < 0) return; // the reference panics on this viewpoint: no frame
    }
    fe::Frame<EMIT> fr(m);
    const size_t o = (size_t)v * (size_t)m.W;
    fr.sc.hor_ocl = s.hor_ocl + o;
    fr.sc.floor_ocl = s.floor_ocl + o;
    fr.sc.ceil_ocl = s.ceil_ocl + o;
    fr.sc.rows[0] = s.rows + 2 * o;
    fr.sc.rows[1] = s.rows + 2 * o + m.W;
    fr.out = out;
    fr.run(views[v], b);
    if (!EMIT && (threadIdx.x & 31u) == 0u) counts[v] = fr.n;
}

cudaError_t launch_frontend(bool emit, const fe::Map &m, const fe::ViewIn *views, const fe::Bases *bases, fe::Counts *counts, int n,
                            const FeScratch &s, const fe::Out &out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int vpb = FE_THREADS / 32; // viewpoints per CTA
    const unsigned blocks = (unsigned)((n + vpb - 1) / vpb);
    if (emit)
        drr_frontend_kernel<true><<<blocks, FE_THREADS, 0, st>>>(m, views, bases, counts, n, s, out);
    else
        drr_frontend_kernel<false><<<blocks, FE_THREADS, 0, st>>>(m, views, bases, counts, n, s, out);
    return cudaGetLastError();
}

} // namespace drr
