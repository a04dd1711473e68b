// issue_peak.cu — measured denominators for the FP32 / MUFU / integer issue rooflines ("measure
// with an FFMA/MUFU micro-benchmark on the box and record it next to HBM").  Stand-alone tool, not part of libhmcgpu.so:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/issue_peak scripts/issue_peak.cu && gpurun_out/issue_peak
//
// Every kernel runs ILP independent dependent-chains per thread, 1024 threads per block, 2 blocks per SM (64 warps per
// SM = full occupancy), long enough (tens of ms) for the clock to settle under load.  It prints one JSON line with
//   ffma   : FP32 FMA warp-instructions/s  (peak = SMs x 4 schedulers x clk)
//   mufu   : MUFU.EX2 warp-instructions/s  (peak = SMs x 4 x clk / 2 ... 16 lanes per clock per SM quarter)
//   imad   : 32-bit IMAD.WIDE (the Philox round: mul.wide.u32) warp-instructions/s, as Philox uses it (each comes with a MOV
//            that widens the 32-bit addend); imad_wide_acc = IMAD.WIDE alone; imad_hi / imad_lo = the two halves issued
//            separately (IMAD.HI.U32, IMAD)
//   mix    : the sweep kernel's rough blend (6 FFMA : 1 MUFU.EX2 : 2 IMAD.WIDE : ~9 IADD3/LOP3/MOV per chain and iteration;
//            the main loop is 147 SASS instructions per 2 iterations with nvcc 12.9 -- kMixInstPerIter, recount with
//            `cuobjdump -sass` after a compiler change)
// `bench.py` divides its measured warp-instructions/s by SMs x 4 x the sampled SM clock; this tool tells how much of that
// nominal issue rate a kernel made of nothing but arithmetic actually reaches at the power-capped clock.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int ILP = 8;
constexpr int THREADS = 1024;

__global__ void __launch_bounds__(THREADS, 2) ffma_kernel(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ float ex2(float v) {
    float r;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__global__ void __launch_bounds__(THREADS, 2) mufu_kernel(float* out, int iters) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = -1.0f - threadIdx.x * 1e-4f - i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = ex2(x[i]);      // stays in (0, 1]: no denormals, no overflow
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(THREADS, 2) imad_kernel(unsigned* out, int iters, unsigned m) {
    unsigned lo[ILP], hi[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { lo[i] = threadIdx.x + i; hi[i] = blockIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {                      // one Philox half-round: 32x32 -> 64 multiply, fold hi into lo
                const unsigned long long p = (unsigned long long)m * lo[i] + hi[i];
                lo[i] = (unsigned)p;
                hi[i] = (unsigned)(p >> 32);
            }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += lo[i] ^ hi[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// IMAD.WIDE alone: the 64-bit accumulator is the addend, so no MOV is needed to widen an operand
__global__ void __launch_bounds__(THREADS, 2) imadw_kernel(unsigned* out, int iters, unsigned m) {
    unsigned long long acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = ((unsigned long long)blockIdx.x << 32) | (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = (unsigned long long)m * (unsigned)acc[i] + acc[i];
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= acc[i];
    if (s == 0x12345678ull) out[blockIdx.x * blockDim.x + threadIdx.x] = (unsigned)s;
}

// the two halves of a Philox product issued separately: IMAD.HI.U32 (mul.hi) and IMAD (mul.lo + add)
__global__ void __launch_bounds__(THREADS, 2) mulhi_kernel(unsigned* out, int iters, unsigned m) {
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = 0x9E3779B9u * (threadIdx.x + 1) + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = __umulhi(x[i], m);                 // values decay to 0: timing is data-independent
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(THREADS, 2) mullo_kernel(unsigned* out, int iters, unsigned m, unsigned c) {
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = x[i] * m + c;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4 chains x (6 FFMA : 1 MUFU.EX2 : 2 IMAD.WIDE : integer folds) per iteration
constexpr double kMixInstPerIter = 73.5;     // counted in the SASS of the main loop (147 per 2 unrolled iterations)
__global__ void __launch_bounds__(THREADS, 2) mix_kernel(float* out, int iters, float a, float b, unsigned m) {
    float x[4], e[4];
    unsigned lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = threadIdx.x * 1e-3f + i; e[i] = -1.f - i; lo[i] = threadIdx.x + i; hi[i] = blockIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = fmaf(x[i], a, b); x[i] = fmaf(x[i], a, e[i]); x[i] = fmaf(x[i], a, b);
            x[i] = fmaf(x[i], b, a); x[i] = fmaf(x[i], a, b); x[i] = fmaf(x[i], b, a);
            e[i] = ex2(e[i]);
            unsigned long long p = (unsigned long long)m * lo[i] + hi[i];
            lo[i] = (unsigned)p; hi[i] = (unsigned)(p >> 32);
            p = (unsigned long long)m * hi[i] + lo[i];
            lo[i] = (unsigned)p ^ 0x9E3779B9u; hi[i] = ((unsigned)(p >> 32) + 0xBB67AE85u) ^ lo[i];
            lo[i] += hi[i];
        }
    }
    float s = 0.f;
    unsigned q = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { s += x[i] + e[i]; q += lo[i] ^ hi[i]; }
    if (s == 12345.678f && q == 77u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <typename F> static int timed(F launch, double* ms_out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();                                                    // warm-up
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *ms_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, blocks = 2 * sms;
    void* out = nullptr;
    CK(cudaMalloc(&out, (size_t)blocks * THREADS * 4));
    const double warps = (double)blocks * THREADS / 32.0;
    double ms_f, ms_m, ms_i, ms_x;
    if (timed([&] { ffma_kernel<<<blocks, THREADS>>>((float*)out, iters, 0.999f, 1e-3f); }, &ms_f)) return 1;
    if (timed([&] { mufu_kernel<<<blocks, THREADS>>>((float*)out, iters); }, &ms_m)) return 1;
    if (timed([&] { imad_kernel<<<blocks, THREADS>>>((unsigned*)out, iters, 0xD2511F53u); }, &ms_i)) return 1;
    if (timed([&] { mix_kernel<<<blocks, THREADS>>>((float*)out, iters, 0.999f, 1e-3f, 0xD2511F53u); }, &ms_x)) return 1;
    double ms_w, ms_h, ms_l;
    if (timed([&] { imadw_kernel<<<blocks, THREADS>>>((unsigned*)out, iters, 0xD2511F53u); }, &ms_w)) return 1;
    if (timed([&] { mulhi_kernel<<<blocks, THREADS>>>((unsigned*)out, iters, 0xD2511F53u); }, &ms_h)) return 1;
    if (timed([&] { mullo_kernel<<<blocks, THREADS>>>((unsigned*)out, iters, 0xD2511F53u, 0x9E3779B9u); }, &ms_l)) return 1;
    const double per_chain = (double)iters * 4 * ILP;            // instructions per warp in the single-op kernels
    const double ffma = warps * per_chain / (ms_f * 1e-3), mufu = warps * per_chain / (ms_m * 1e-3),
                 imad = warps * per_chain / (ms_i * 1e-3), mix = warps * (double)iters * kMixInstPerIter / (ms_x * 1e-3);
    printf("{\"tool\": \"issue_peak\", \"gpu\": \"%s\", \"sms\": %d, \"iters\": %d, "
           "\"ffma_gwarp_inst_s\": %.1f, \"mufu_ex2_gwarp_inst_s\": %.1f, \"imad_wide_gwarp_inst_s\": %.1f, \"mix_gwarp_inst_s\": %.1f, "
           "\"imad_wide_acc_gwarp_inst_s\": %.1f, \"imad_hi_gwarp_inst_s\": %.1f, \"imad_lo_gwarp_inst_s\": %.1f, "
           "\"ffma_lane_ops_s\": %.4g, \"mufu_lane_ops_s\": %.4g, "
           "\"ms\": {\"ffma\": %.3f, \"mufu\": %.3f, \"imad\": %.3f, \"mix\": %.3f}, "
           "\"nominal_issue_gwarp_inst_s_at_max_clock\": %.1f}\n",
           prop.name, sms, iters, ffma / 1e9, mufu / 1e9, imad / 1e9, mix / 1e9,
           warps * per_chain / (ms_w * 1e-3) / 1e9, warps * per_chain / (ms_h * 1e-3) / 1e9, warps * per_chain / (ms_l * 1e-3) / 1e9,
           ffma * 32, mufu * 32, ms_f, ms_m, ms_i, ms_x,
           sms * 4.0 * prop.clockRate * 1e3 / 1e9);
    cudaFree(out);
    return 0;
}
