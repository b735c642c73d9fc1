// ubench.cu -- instruction-throughput microbenchmarks for sm_100a (B200), used to build the per-pixel cost model of the
// draw kernels.  Each kernel runs CHAINS independent dependency chains of one instruction per
// thread, ITERS times; the result is warp-instructions per cycle per SM (4 = one per scheduler per cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

#define KERNEL(name, decl, init, body, fin)                                                     \
    __global__ void __launch_bounds__(256) name(uint32_t *out, uint32_t seed) {                  \
        decl;                                                                                   \
        init;                                                                                   \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) {                                \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { body; }                        \
        }                                                                                       \
        uint32_t acc = 0;                                                                       \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { fin; }                             \
        if (acc == 0x12345678u) out[threadIdx.x] = acc;                                         \
    }

#define F_DECL float x[CHAINS]; float k1 = __uint_as_float(seed | 0x3f800000u), k2 = k1 * 0.5f
#define F_INIT for (int c = 0; c < CHAINS; ++c) x[c] = (float)(threadIdx.x + c) * 1e-3f
#define F_FIN acc += __float_as_uint(x[c])
#define U_DECL uint32_t x[CHAINS]; uint32_t k1 = seed | 1u, k2 = seed * 7u + 3u
#define U_INIT for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 17u + c
#define U_FIN acc += x[c]
#define D_DECL float2 x[CHAINS]; float k1 = __uint_as_float(seed | 0x3f800000u), k2 = k1 * 0.5f
#define D_INIT for (int c = 0; c < CHAINS; ++c) x[c] = make_float2((float)(threadIdx.x + c) * 1e-3f, (float)c)
#define D_FIN acc += __float_as_uint(x[c].x) + __float_as_uint(x[c].y)

KERNEL(k_ffma, F_DECL, F_INIT, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(k1), "f"(k2)), F_FIN)
KERNEL(k_fmul, F_DECL, F_INIT, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k1)), F_FIN)
KERNEL(k_fadd, F_DECL, F_INIT, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k1)), F_FIN)
KERNEL(k_fadd_rz, F_DECL, F_INIT, asm volatile("add.rz.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k1)), F_FIN)
KERNEL(k_ffma2, D_DECL, D_INIT,
       asm volatile("{.reg .b64 a, b, d; mov.b64 a, {%0,%1}; mov.b64 b, {%2,%2}; mov.b64 d, {%3,%3}; fma.rn.f32x2 a, a, b, d; mov.b64 {%0,%1}, a;}"
                    : "+f"(x[c].x), "+f"(x[c].y) : "f"(k1), "f"(k2)), D_FIN)
KERNEL(k_fmul2, D_DECL, D_INIT,
       asm volatile("{.reg .b64 a, b; mov.b64 a, {%0,%1}; mov.b64 b, {%2,%2}; mul.rn.f32x2 a, a, b; mov.b64 {%0,%1}, a;}"
                    : "+f"(x[c].x), "+f"(x[c].y) : "f"(k1)), D_FIN)
KERNEL(k_fadd2_rz, D_DECL, D_INIT,
       asm volatile("{.reg .b64 a, b; mov.b64 a, {%0,%1}; mov.b64 b, {%2,%2}; add.rz.f32x2 a, a, b; mov.b64 {%0,%1}, a;}"
                    : "+f"(x[c].x), "+f"(x[c].y) : "f"(k1)), D_FIN)
KERNEL(k_f2i_s16, F_DECL, F_INIT, { int r; asm volatile("cvt.rzi.s16.f32 %0, %1;" : "=r"(r) : "f"(x[c])); x[c] = __int_as_float(r | 0x3f000000); }, F_FIN)
KERNEL(k_f2i_s32, F_DECL, F_INIT, { int r; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(r) : "f"(x[c])); x[c] = __int_as_float(r | 0x3f000000); }, F_FIN)
KERNEL(k_i2f, U_DECL, U_INIT, { float r; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(r) : "r"(x[c])); x[c] = __float_as_uint(r); }, U_FIN)
KERNEL(k_rcp, F_DECL, F_INIT, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c])), F_FIN)
KERNEL(k_iadd, U_DECL, U_INIT, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1)), U_FIN)
KERNEL(k_lop3, U_DECL, U_INIT, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2)), U_FIN)
KERNEL(k_prmt, U_DECL, U_INIT, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2 & 0x7777u)), U_FIN)
KERNEL(k_imad, U_DECL, U_INIT, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2)), U_FIN)
KERNEL(k_imadhi, U_DECL, U_INIT, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1)), U_FIN)
KERNEL(k_imadwide, U_DECL; unsigned long long w[CHAINS], U_INIT; for (int c = 0; c < CHAINS; ++c) w[c] = c,
       asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(k1)), acc += x[c] + (uint32_t)w[c] + (uint32_t)(w[c] >> 32))
KERNEL(k_shl_add, U_DECL, U_INIT, asm volatile("{.reg .u32 t; shl.b32 t, %0, 4; add.u32 %0, t, %1;}" : "+r"(x[c]) : "r"(k1)), U_FIN)
KERNEL(k_fmnmx, F_DECL, F_INIT, asm volatile("max.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k1)), F_FIN)
KERNEL(k_shfl, U_DECL, U_INIT, asm volatile("shfl.sync.idx.b32 %0, %0, %1, 31, 0xffffffff;" : "+r"(x[c]) : "r"(k1 & 31u)), U_FIN)
// mixes: alternate an FMA-pipe and an ALU-pipe instruction
KERNEL(k_mix_ffma_iadd, F_DECL; uint32_t u[CHAINS], F_INIT; for (int c = 0; c < CHAINS; ++c) u[c] = c,
       { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(k1), "f"(k2)); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(seed)); },
       acc += __float_as_uint(x[c]) + u[c])
KERNEL(k_mix_ffma2_iadd, D_DECL; uint32_t u[CHAINS], D_INIT; for (int c = 0; c < CHAINS; ++c) u[c] = c,
       { asm volatile("{.reg .b64 a, b, d; mov.b64 a, {%0,%1}; mov.b64 b, {%2,%2}; mov.b64 d, {%3,%3}; fma.rn.f32x2 a, a, b, d; mov.b64 {%0,%1}, a;}"
                      : "+f"(x[c].x), "+f"(x[c].y) : "f"(k1), "f"(k2));
         asm volatile("add.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(seed)); },
       acc += __float_as_uint(x[c].x) + __float_as_uint(x[c].y) + u[c])
KERNEL(k_mix_ffma_f2i, F_DECL; int u[CHAINS], F_INIT; for (int c = 0; c < CHAINS; ++c) u[c] = c,
       { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(k1), "f"(k2)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(k2), "f"(k1));
         asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(k1), "f"(k2)); asm volatile("cvt.rzi.s16.f32 %0, %1;" : "=r"(u[c]) : "f"(x[c])); },
       acc += __float_as_uint(x[c]) + (uint32_t)u[c])

// shared memory: 128-bit loads at data-dependent (conflicting) addresses, 32-bit loads, 32-bit stores
__global__ void __launch_bounds__(256) k_lds128(uint32_t *out, uint32_t seed) {
    __shared__ float4 tab[256];
    tab[threadIdx.x] = make_float4(threadIdx.x, 1, 2, __uint_as_float((threadIdx.x * 37u + seed) & 255u));
    __syncthreads();
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = (threadIdx.x * 7u + c * 13u) & 255u;
    float s = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            float4 v = tab[idx[c]];
            idx[c] = __float_as_uint(v.w);
            s += v.x;
        }
    }
    if (s == 1.2345f) out[threadIdx.x] = idx[0];
}
__global__ void __launch_bounds__(256) k_lds32(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t tab[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) tab[i] = (i * 37u + seed) & 1023u;
    __syncthreads();
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = (threadIdx.x + c * 32u) & 1023u;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) idx[c] = tab[idx[c]];
    }
    uint32_t acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc += idx[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_sts32(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t tab[256 * CHAINS];
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tab + c * 256 + threadIdx.x)), "r"(seed + it) : "memory");
    }
    __syncthreads();
    if (tab[threadIdx.x] == 0x12345678u) out[threadIdx.x] = 1;
}
// L1-hit global loads of u16 at scattered addresses within a 64 KB window (texel fetch pattern)
__global__ void __launch_bounds__(256) k_ldg16(uint32_t *out, uint32_t seed, const uint16_t *tex) {
    uint32_t idx[CHAINS];
    for (int c = 0; c < CHAINS; ++c) idx[c] = (threadIdx.x * 2u + c * 977u) & 32767u;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) idx[c] = (__ldg(tex + idx[c]) + seed) & 32767u;
    }
    uint32_t acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc += idx[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <class F>
static void run(const char *name, F launch, double ops_per_iter_per_thread) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    launch(blocks);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(blocks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_inst = (double)blocks * 8 * ITERS * CHAINS * ops_per_iter_per_thread;
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM (at %d MHz nominal)  err=%s\n", name, ms, warp_inst / cycles / 148.0, clk_khz / 1000,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *out;
    cudaMalloc(&out, 4096);
    uint16_t *tex;
    cudaMalloc(&tex, 65536);
    cudaMemset(tex, 1, 65536);
    const uint32_t seed = 12345;
#define RUN(k, n) run(#k, [&](int b) { k<<<b, 256>>>(out, seed); }, n)
    for (int rep = 0; rep < 2; ++rep) {
        RUN(k_ffma, 1); RUN(k_fmul, 1); RUN(k_fadd, 1); RUN(k_fadd_rz, 1); RUN(k_ffma2, 1); RUN(k_fmul2, 1); RUN(k_fadd2_rz, 1);
        RUN(k_f2i_s16, 1); RUN(k_f2i_s32, 1); RUN(k_i2f, 1); RUN(k_rcp, 1); RUN(k_iadd, 1); RUN(k_lop3, 1); RUN(k_prmt, 1); RUN(k_imad, 1);
        RUN(k_imadhi, 1); RUN(k_imadwide, 1); RUN(k_shl_add, 1); RUN(k_fmnmx, 1); RUN(k_shfl, 1);
        RUN(k_mix_ffma_iadd, 2); RUN(k_mix_ffma2_iadd, 2); RUN(k_mix_ffma_f2i, 4);
        RUN(k_lds128, 1); RUN(k_lds32, 1); RUN(k_sts32, 1);
        run("k_ldg16", [&](int b) { k_ldg16<<<b, 256>>>(out, seed, tex); }, 1);
        printf("----\n");
    }
    return 0;
}
