// ubench_tex.cu -- does the texture path (TEX pipe) run beside the LSU pipe?  Random 16-byte palette lookups through
// shared memory (LDS.128), through a 1D linear texture (tex1Dfetch<float4>), and both in the same loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tex ubench_tex.cu && ./ubench_tex
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define CHAINS 4
__global__ void __launch_bounds__(256) k_lds(uint32_t *out, const float4 *pal, int coherent) {
    __shared__ float4 tab[256];
    tab[threadIdx.x] = pal[threadIdx.x];
    __syncthreads();
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = ((coherent ? (threadIdx.x >> 2) : threadIdx.x) * 7u + c * 13u) & 255u;
    float s = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { float4 v = tab[idx[c]]; idx[c] = __float_as_uint(v.w) & 255u; s += v.x; }
    }
    if (s == 1.2345f) out[threadIdx.x] = idx[0];
}
__global__ void __launch_bounds__(256) k_tex(uint32_t *out, cudaTextureObject_t tex, int coherent) {
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = ((coherent ? (threadIdx.x >> 2) : threadIdx.x) * 7u + c * 13u) & 255u;
    float s = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { float4 v = tex1Dfetch<float4>(tex, (int)idx[c]); idx[c] = __float_as_uint(v.w) & 255u; s += v.x; }
    }
    if (s == 1.2345f) out[threadIdx.x] = idx[0];
}
__global__ void __launch_bounds__(256) k_both(uint32_t *out, const float4 *pal, cudaTextureObject_t tex, int coherent) {
    __shared__ float4 tab[256];
    tab[threadIdx.x] = pal[threadIdx.x];
    __syncthreads();
    uint32_t idx[CHAINS], jdx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { idx[c] = ((coherent ? (threadIdx.x >> 2) : threadIdx.x) * 7u + c * 13u) & 255u; jdx[c] = (idx[c] * 5u + 1u) & 255u; }
    float s = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            float4 v = tab[idx[c]]; idx[c] = __float_as_uint(v.w) & 255u; s += v.x;
            float4 w = tex1Dfetch<float4>(tex, (int)jdx[c]); jdx[c] = __float_as_uint(w.w) & 255u; s += w.y;
        }
    }
    if (s == 1.2345f) out[threadIdx.x] = idx[0] + jdx[0];
}
__global__ void __launch_bounds__(256) k_tex16(uint32_t *out, cudaTextureObject_t tex) {
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = (threadIdx.x * 2u + c * 977u) & 32767u;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) idx[c] = (tex1Dfetch<unsigned short>(tex, (int)idx[c]) + 3u * threadIdx.x) & 32767u;
    }
    uint32_t acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc += idx[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
template <class F> static void run(const char *name, F launch, double ops) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    launch(blocks); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(blocks); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double wi = (double)blocks * 8 * ITERS * CHAINS * ops, cyc = ms * 1e-3 * khz * 1e3;
    printf("%-22s %8.3f ms  %6.3f warp-lookups/clk/SM = %5.2f clk per warp-lookup  err=%s\n", name, ms, wi / cyc / 148.0, cyc * 148.0 / wi, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint32_t *out; cudaMalloc(&out, 4096);
    float4 h[256];
    for (int i = 0; i < 256; ++i) { uint32_t nx = (i * 37u + 11u) & 255u; h[i] = make_float4((float)i, 1.f, 2.f, 0.f); memcpy(&h[i].w, &nx, 4); }
    float4 *pal; cudaMalloc(&pal, sizeof(h)); cudaMemcpy(pal, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = pal; rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = sizeof(h);
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    uint16_t *t16; cudaMalloc(&t16, 65536); cudaMemset(t16, 1, 65536);
    cudaResourceDesc rd2 = {}; rd2.resType = cudaResourceTypeLinear; rd2.res.linear.devPtr = t16; rd2.res.linear.desc = cudaCreateChannelDesc<unsigned short>(); rd2.res.linear.sizeInBytes = 65536;
    cudaTextureObject_t tex16; cudaCreateTextureObject(&tex16, &rd2, &td, nullptr);
    for (int rep = 0; rep < 2; ++rep)
        for (int coh = 0; coh < 2; ++coh) {
            printf("-- %s indices\n", coh ? "4-lane coherent" : "random");
            run("lds128", [&](int b) { k_lds<<<b, 256>>>(out, pal, coh); }, 1);
            run("tex float4", [&](int b) { k_tex<<<b, 256>>>(out, tex, coh); }, 1);
            run("lds128 + tex (pairs)", [&](int b) { k_both<<<b, 256>>>(out, pal, tex, coh); }, 1);
            run("tex u16 (64 KB)", [&](int b) { k_tex16<<<b, 256>>>(out, tex16); }, 1);
        }
    return 0;
}
