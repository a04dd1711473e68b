// rng.cuh — counter-based random streams for the Gibbs path.
//
// Replaces the reference's global MersenneTwister + Distributions.jl samplers (src/Hmc.jl:320 InverseGamma,
// :334 Normal, :355/:367 Dirichlet, :464/:481 Categorical).  Every random number is a pure function of
// (seed, chain id, sweep, purpose, index): Philox4x32 with key = seed and counter = {block, purpose, sweep, chain},
// so results do not depend on scheduling or on how chains are sharded over GPUs.  The parameter draws use the 10-round
// generator; the STATES stream (one uniform per time step of every sweep: a third of the backward pass's instructions)
// uses Philox4x32-7, the fewest rounds Random123 reports as passing BigCrush ("crush-resistant").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hmc {

enum : uint32_t { KIND_STATES = 0u, KIND_MU = 1u, KIND_SIGMA = 2u, KIND_RHO = 3u, KIND_A = 4u };

struct RngKey {
    uint32_t k0, k1;   // seed lo / hi
    uint32_t chain;    // global chain id = window_id * n_chains + chain
};

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}
constexpr int kStateRounds = 7;
// parameter draws (KIND_MU / SIGMA / RHO / A)
__device__ __forceinline__ uint4 rng_block(const RngKey& k, uint32_t sweep, uint32_t purpose, uint32_t block) {
    return philox4x32<10>(block, purpose, sweep, k.chain, k.k0, k.k1);
}
// backward-sampler uniforms (KIND_STATES): block b holds the uniforms consumed 4b .. 4b+3
__device__ __forceinline__ uint4 rng_block_states(const RngKey& k, uint32_t sweep, uint32_t block) {
    return philox4x32<kStateRounds>(block, (KIND_STATES << 16), sweep, k.chain, k.k0, k.k1);
}

// uniform in (0,1): fp64 = (w + 0.5) * 2^-32 exactly
template <typename R> __device__ __forceinline__ R u01(uint32_t w);
template <> __device__ __forceinline__ double u01<double>(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
// fp32: the word rounded toward zero to 24 significant bits (one I2FP), offset by 2^-33 so that u > 0; u < 1 always
template <> __device__ __forceinline__ float u01<float>(uint32_t w) { return fmaf(__uint2float_rz(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

template <typename R> struct M;
template <> struct M<double> {
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double cos2pi(double u) { return ::cos(6.283185307179586 * u); }
    static __device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
};
// fp32 draws use the SFU approximations (lg2/ex2/sin-cos/rsqrt, ~1e-6 relative): the conjugate draws are ~10 % of the
// sweep kernel's instructions with the libm versions, and draw-level accuracy is far inside the fp32 path's tolerance.
template <> struct M<float> {
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
    static __device__ __forceinline__ float sqrt(float x) { return __fsqrt_rn(x); }
    static __device__ __forceinline__ float cos2pi(float u) { return __cosf(6.283185307179586f * u); }
    static __device__ __forceinline__ float pow(float x, float y) { return __powf(x, y); }
};

// Box-Muller, cos branch, from two words
template <typename R> __device__ __forceinline__ R normal_from(uint32_t w0, uint32_t w1) {
    const R r = M<R>::sqrt(R(-2) * M<R>::log(u01<R>(w0)));
    return r * M<R>::cos2pi(u01<R>(w1));
}

// Marsaglia & Tsang gamma(shape, 1); attempt k = block k of the draw's own purpose stream:
// words 0,1 -> normal, word 2 -> acceptance uniform, word 3 -> boost uniform (shape < 1).
template <typename R> __device__ __noinline__ R gamma_mt(R shape, const RngKey key, uint32_t sweep, uint32_t purpose) {
    const R a = shape < R(1) ? shape + R(1) : shape;
    const R d = a - R(1) / R(3), c = R(1) / M<R>::sqrt(R(9) * d);
    R g = d;
    for (uint32_t k = 0; k < 4096u; ++k) {
        const uint4 w = rng_block(key, sweep, purpose, k);
        const R z = normal_from<R>(w.x, w.y);
        R v = R(1) + c * z;
        if (v <= R(0)) continue;
        v = v * v * v;
        const R u = u01<R>(w.z);
        const R z2 = z * z;
        if (u < R(1) - R(0.0331) * z2 * z2 || M<R>::log(u) < R(0.5) * z2 + d * (R(1) - v + M<R>::log(v))) {
            g = d * v;
            if (shape < R(1)) g *= M<R>::pow(u01<R>(w.w), R(1) / shape);
            break;
        }
    }
    return g;
}

}  // namespace hmc
