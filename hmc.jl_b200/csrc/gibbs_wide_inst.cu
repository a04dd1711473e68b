// gibbs_wide_inst.cu — instantiation of the lane-per-state sweep kernel (K = 5..32) for one precision:
// compiled with -DHMC_R=<float|double>.
#include "gibbs_wide_kernel.cuh"

namespace hmc {

template <typename R, int W>
static cudaError_t launch_wide_w(const GibbsLaunch& cfg, const GibbsArgs& a, int K, const long long* slot_pi_off, cudaStream_t st) {
    const int per_block = kWideThreads / W;
    const unsigned grid = (unsigned)((a.n_slots + per_block - 1) / per_block);
    const size_t smem = per_block * wide_group_bytes<R>(K, W);
    auto go = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, kWideThreads, smem, st>>>(a, K, slot_pi_off);
        return cudaGetLastError();
    };
    if (cfg.flags & 16u /*HMCGPU_FLAG_LOGLIK*/) return go(gibbs_wide_kernel<R, W, true>);
    return go(gibbs_wide_kernel<R, W, false>);
}

template <typename R> cudaError_t launch_gibbs_wide(const GibbsLaunch& cfg, const GibbsArgs& a, int K, const long long* slot_pi_off, cudaStream_t st) {
    if (K <= 8) return launch_wide_w<R, 8>(cfg, a, K, slot_pi_off, st);
    if (K <= 16) return launch_wide_w<R, 16>(cfg, a, K, slot_pi_off, st);
    return launch_wide_w<R, 32>(cfg, a, K, slot_pi_off, st);
}

template cudaError_t launch_gibbs_wide<HMC_R>(const GibbsLaunch&, const GibbsArgs&, int, const long long*, cudaStream_t);

}  // namespace hmc
