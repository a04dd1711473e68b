// gibbs_kernel.cuh — the fused Gibbs sweep kernel: one thread owns one (window x chain) Markov chain and runs
// whole sweeps on-device (src/Hmc.jl:486-515 gibbssweep!, looped as in gibbssample! :545-560):
//   conjugate draws -> forward filter (pif spilled to HBM, coalesced [row][k][lane]) -> relabel/emit the draw ->
//   backward state sampling fused with the next sweep's sufficient statistics (X never touches memory) ->
//   optional backward smoother accumulation.
//
// Mapping.  A warp = 32 chains in lockstep ("warp task"); lanes of a warp share one contiguous pif tile so every
// global access is a full 128-byte (fp32) / 256-byte (fp64) line.  The cost of a warp task is its window length, which
// varies 6x across the expanding windows, and the SM warp arbiter is priority-based (starved warps finish late), so
// balance comes from the launch structure (hmcgpu.cu): tasks are sorted longest-first, split into interleaved groups,
// and each group runs short launches (a few sweeps) on its own stream — while one group's launch drains, the block
// scheduler backfills the freed SM slots with the pending blocks of the other groups' launches.  Windows inside a warp are right-aligned in time: loop row j
// holds t = j - (T_warp - T_lane), which keeps the first backward step and every Philox block warp-uniform.
#pragma once
#include <type_traits>
#include "hmm_device.cuh"

namespace hmc {

// experiment knobs (defaults = the measured best; see profiles/ for the A/B runs)
#ifndef HMC_SEL_LDS
#define HMC_SEL_LDS 1      // 1: A[:,x] and the counter increment come from a per-thread shared-memory table; 0: register selects
#endif

#ifndef HMC_BACK_INLINE
#define HMC_BACK_INLINE 0  // experiment: backward pass inlined into the sweep loop (no R2UR descriptor moves; shares the kernel's register allocation)
#endif
#if HMC_BACK_INLINE
#define HMC_BACK_ATTR __forceinline__
#else
#define HMC_BACK_ATTR __noinline__
#endif
#ifndef HMC_FWD_INLINE
#define HMC_FWD_INLINE 0
#endif
#if HMC_FWD_INLINE
#define HMC_FWD_ATTR __forceinline__
#else
#define HMC_FWD_ATTR __noinline__
#endif

#ifndef HMC_ASYNC
#define HMC_ASYNC 1        // 1: backward rows are staged through a per-warp shared-memory ring with cp.async (LDGSTS)
#endif
#ifndef HMC_TMA
#define HMC_TMA 0          // 1: the ring is filled by ONE 1-D bulk copy per group (cp.async.bulk, the TMA engine; SASS UBLKCP) issued by an
#endif                     //    elected lane and completed on an mbarrier, instead of K*sizeof(R)/4 LDGSTS per lane (K <= 4 kernels)
#ifndef HMC_RING_STAGES
#define HMC_RING_STAGES 4  // groups of 4 rows in flight per warp
#endif

constexpr int kMaxH = 16;
constexpr int kRing = HMC_RING_STAGES;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
// same with the 32-bit shared-window address computed once by the caller (the generic -> shared conversion reads a special
// register and is not hoisted out of loops by the compiler)
__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// ---- TMA 1-D bulk copy + mbarrier (HMC_TMA).  All waits are bounded: a lost completion turns into an error count, never a hang.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_inval(unsigned bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
constexpr int kTmaWaitSpins = 1 << 16;                               // try_wait sleeps in hardware between polls: far beyond any real latency
static __device__ __noinline__ bool mbar_wait_bounded(unsigned bar, unsigned parity) {     // the cold path of a wait: false = the phase never completed
    for (int spins = 0; spins < kTmaWaitSpins; ++spins)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ bool elect_one() {                        // one lane of the (converged) warp
    unsigned ok;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(ok));
    return ok != 0u;
}

// The passes are out-of-line functions, where pointers arriving through their by-value arguments would be generic
// (LD.E/ST.E).  Shared memory is therefore always re-derived from the kernel's dynamic shared array, and global memory
// goes through the explicit-space intrinsics below (LDG/STG in SASS).
__device__ __forceinline__ unsigned char* smem_base() {
    extern __shared__ __align__(16) unsigned char hmc_smem[];
    return hmc_smem;
}
template <typename T> __device__ __forceinline__ T ld_stream(const T* p) { return __ldcg(p); }     // pif rows: L2 only
template <typename T> __device__ __forceinline__ T ld_ro(const T* p) { return __ldg(p); }          // y: read-only, L1
template <typename T> __device__ __forceinline__ void st_stream(T* p, T v) { __stcg(p, v); }
// Where a pass reads its observations from: global memory (read-only path, L1) or — the windows of ONE short series, the headline
// case — a copy of the series in shared memory, addressed through the 32-bit shared window (no 64-bit pointer arithmetic, no
// memory descriptor to move into a uniform register in front of every load).  Pointer-like: + / - in elements, ld() reads.
template <typename T, bool SM> struct YRef;
template <typename T> struct YRef<T, false> {
    const T* p;
    __device__ __forceinline__ YRef operator+(long long d) const { return YRef{p + d}; }
    __device__ __forceinline__ YRef operator-(long long d) const { return YRef{p - d}; }
    __device__ __forceinline__ YRef& operator+=(long long d) { p += d; return *this; }
    __device__ __forceinline__ YRef& operator-=(long long d) { p -= d; return *this; }
    __device__ __forceinline__ T ld() const { return __ldg(p); }
    __device__ __forceinline__ const void* gptr() const { return p; }
};
template <typename T> struct YRef<T, true> {
    unsigned a;                                                       // byte address in the shared window
    __device__ __forceinline__ YRef operator+(long long d) const { return YRef{a + (unsigned)((int)d * (int)sizeof(T))}; }
    __device__ __forceinline__ YRef operator-(long long d) const { return YRef{a - (unsigned)((int)d * (int)sizeof(T))}; }
    __device__ __forceinline__ YRef& operator+=(long long d) { a += (unsigned)((int)d * (int)sizeof(T)); return *this; }
    __device__ __forceinline__ YRef& operator-=(long long d) { a -= (unsigned)((int)d * (int)sizeof(T)); return *this; }
    __device__ __forceinline__ T ld() const {
        T v;
        if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    __device__ __forceinline__ const void* gptr() const { return nullptr; }
};
// observations of one series kept in shared memory by the thread-per-chain kernel (GibbsArgs::y_sm_elems): at most this many bytes
#ifndef HMC_YSM_F32
#define HMC_YSM_F32 0          // 1: fp32 also reads the one shared series from the shared-memory copy (A/B knob)
#endif
template <typename R> __host__ __device__ constexpr bool y_in_smem() { return sizeof(R) == 8 || HMC_YSM_F32; }
constexpr int kYSmemMaxBytes = 8 * 1024;    // (1024 fp64 observations; tables + rings + forecasts are ~61 KB per block at 3 blocks per SM)
// Batches of distinct series (yld > 1) stream y from HBM once per pass: the rows are pulled into L1 kYAhead steps ahead
// of their use so the DRAM latency does not sit in front of every group of 4 steps.  (Windows of ONE series keep y in L1.)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef HMC_YAHEAD
#define HMC_YAHEAD 16
#endif
constexpr int kYAhead = HMC_YAHEAD;
// The pif spill of a warp task is tiled by groups of 4 time steps: tile q holds rows 4q..4q+3 (rows right-aligned to a
// multiple of 4: row j of the warp frame sits at j + pad, pad = (4 - Tw % 4) % 4), element (state k, lane, row r) at
// (k*32 + lane)*4 + r.  A lane's 4 rows of one state are 16 (fp32) / 32 (fp64) contiguous bytes: the forward pass
// stores them with one (two) 128-bit STG per state and group, the backward pass reads them back from the cp.async ring
// with one (two) LDS.128 — instead of 4 scalar accesses each.
__device__ __forceinline__ void st_quad(float* p, float a, float b, float c, float d) { __stcg(reinterpret_cast<float4*>(p), make_float4(a, b, c, d)); }
__device__ __forceinline__ void st_quad(double* p, double a, double b, double c, double d) {
    __stcg(reinterpret_cast<double2*>(p), make_double2(a, b));
    __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(c, d));
}
__device__ __forceinline__ void ld_quad_shared(const float* p, float& a, float& b, float& c, float& d) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void ld_quad_shared(const double* p, double& a, double& b, double& c, double& d) {
    const double2 v = *reinterpret_cast<const double2*>(p), w = *(reinterpret_cast<const double2*>(p) + 1);
    a = v.x; b = v.y; c = w.x; d = w.y;
}
// the same through a 32-bit shared-window address (ld.shared.v4 / 2 x ld.shared.v2.f64): no generic-to-shared base in the loop
#ifndef HMC_RING_LDS32
#define HMC_RING_LDS32 1
#endif
__device__ __forceinline__ void lds_quad(unsigned saddr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(saddr));
}
__device__ __forceinline__ void lds_quad(unsigned saddr, double& a, double& b, double& c, double& d) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(saddr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(d) : "r"(saddr + 16u));
}
// a 16-byte-aligned struct read from a 32-bit shared-window address (ld.shared.v4 per 16 bytes)
template <typename T> __device__ __forceinline__ void lds_struct(unsigned saddr, T& out) {
    static_assert(sizeof(T) % 16 == 0 && alignof(T) >= 16, "16-byte granules");
    uint4 w[sizeof(T) / 16];
#pragma unroll
    for (int q = 0; q < (int)(sizeof(T) / 16); ++q)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[q].x), "=r"(w[q].y), "=r"(w[q].z), "=r"(w[q].w) : "r"(saddr + 16u * q));
    memcpy(&out, w, sizeof(T));
}
constexpr int kGibbsThreads = 128;
#ifndef HMC_MINBLOCKS
#define HMC_MINBLOCKS 4
#endif
constexpr int kGibbsMinBlocks = HMC_MINBLOCKS;   // 4 blocks x 128 threads per SM -> 128 registers, 16 resident warps (5 blocks / 96 registers was the best before the tiled spill layout; measured again after it: 4 is +5 %)
// K = 5..8 keep K x K matrices per thread: 2 blocks per SM (shared-memory tables and rings allow no more), 255 registers
#ifndef HMC_MINBLOCKS_F64
#define HMC_MINBLOCKS_F64 4
#endif
#ifndef HMC_MINBLOCKS_K56
#define HMC_MINBLOCKS_K56 3
#endif
#ifndef HMC_MINBLOCKS_SIG
#define HMC_MINBLOCKS_SIG 4
#endif
// fp64 state needs twice the registers: 4 blocks per SM at 128 registers with a depth-3 ring (7.27e10 on C2) or 3 at 168 with depth 4
// (7.06e10); 2 blocks at 255 registers 5.6e10, 96 registers spill (scripts/r2_call_h.sh)
template <typename R, int K, bool SIG> constexpr int gibbs_min_blocks() {
    return K > 6 ? 2 : (K > 4 ? HMC_MINBLOCKS_K56 : (sizeof(R) == 8 ? HMC_MINBLOCKS_F64 : (SIG ? HMC_MINBLOCKS_SIG : kGibbsMinBlocks)));
}
// cp.async ring depth: the ring of a warp is stages x 4 rows x K x 32 lanes
// (fp64, K <= 4: depth 3 — with depth 4 the shared memory of a block allows only 3 blocks per SM; 4 blocks at depth 3 measured +3 %)
#ifndef HMC_RING_STAGES_F64
#define HMC_RING_STAGES_F64 3
#endif
template <typename R, int K> __host__ __device__ constexpr int gibbs_ring_stages() { return K <= 4 ? (sizeof(R) == 8 ? HMC_RING_STAGES_F64 : kRing) : (sizeof(R) == 4 ? 3 : 2); }

struct GibbsArgs {
    int n_slots;                 // chains incl. padding, multiple of 32
    int n_warps;                 // warp tasks = n_slots/32
    // this launch covers the warp tasks task0, task0 + task_stride, ... (n_tasks of them): one task group
    int task0, task_stride, n_tasks;
    const int* T;                // [n_slots] window length (0 = padding lane)
    const long long* ybase;      // [n_slots] element offset of the window's first observation in y
    int yld;                     // stride between consecutive time steps in y (= n_series, time-major)
    const void* y;               // R*
    const int* warp_T;           // [n_warps] max T within the warp
    const long long* warp_pi_off;// [n_warps] element offset of the warp's pif tile
    void* pi;                    // R*  pif spill: tile[(row*K + k)*32 + lane]
    void* pib_acc;               // R*  same layout, smoothed-probability sums of this chunk (SMOOTH only)
    void* fc_acc;                // R*  in-sample forecast sums of this chunk: tile[(row*n_h + j)*32 + lane] at offset warp_pi_off/K*n_h
    // chain state between launches
    int* cnt;                    // [K][n_slots]
    int* trans;                  // [K*K][n_slots]
    void* Sd;                    // R* [K][n_slots]  sum (y-c)
    void* Qd;                    // R* [K][n_slots]  sum (y-c)^2
    int* events;                 // [n_slots]
    const void* cshift;          // R* [n_slots]
    const void* totS;            // R* [n_slots]  sum over the window of (y-c)
    const void* totQ;            // R* [n_slots]  sum over the window of (y-c)^2
    const void* xi;              // R* [K][n_slots]
    double alpha[32], nu[32], beta0[32], beta[32];
    unsigned k0, k1;
    const unsigned* chain_id;    // [n_slots]
    long long sweep0;            // global index of the first sweep of this launch
    int n_sweeps;
    long long burnin;
    // per-draw outputs of this chunk: out[(f*chunk + i)*n_slots + slot], i = draw index within the chunk
    void* out;
    int chunk;
    long long draw0;             // global draw index of chunk slot 0
    int n_h;
    int h_sorted[kMaxH];         // horizons ascending
    int h_slot[kMaxH];           // original position of each sorted horizon
    const void* yfut;            // R* [n_h][n_slots] realised y at end+h (NaN when outside the series)
    unsigned flags;
    // ---- signals tier (SIG kernels only; estimatesignals! src/Hmc.jl:868-914, mask rules :267-300 and :380-383)
    const void* sigw;            // R* emission z-scale, time-major [y_len][sld]: +1 = observation, -1/(1+kappa) = signal
    const long long* sbase;      // [n_slots] element offset of the window's first time step in sigw
    int sld;                     // stride between consecutive time steps in sigw (1 = one mask for every series, else n_series)
    int* cntM;                   // [K][n_slots] signals per state (cnt holds the observations only)
    void* Sm;                    // R* [K][n_slots]  sum (y-c) over the signals of each state (Sd/Qd: observations only)
    void* Qm;                    // R* [K][n_slots]  sum (y-c)^2 over the signals
    const void* totSm;           // R* [n_slots]  window totals over the signals (totS/totQ: over the observations)
    const void* totQm;
    const int* totM;             // [n_slots] number of signals in the window
    double kappa;
    int pi_back;                 // the pi_end field of a draw is the smoothed marginal pib[N - pi_back, :] (:893); 0 = last row
    int y_sm_elems;              // > 0: ONE series of this many observations (yld = 1), copied into shared memory by every block of the thread-per-chain kernel
    int all_signal;              // every time step of every series is a signal (the 'make everything a signal' runs, code/run_hmm.jl:160-176)
    // ---- segment kernel (gibbs_seg_kernel.cuh)
    int seg_warm;                // warm-up time steps in front of a segment
    int seg_barriers;            // block-wide phase barriers per sweep (0, 1: before the filter, 2: also before the sampler)
    unsigned long long* diag;    // [2] chain-sweeps whose warm-up failed the verification (exact products used instead); groups of 4 steps walked speculatively, summed over chain-sweeps; may be NULL
};

// Transition counts n_ij of the sampled path, packed: one word per origin state, one bit-field per destination.
// K > 4: always 64-bit rows; the fields (64/K bits, 8 at K = 8) are flushed into plain counters before they can overflow.
template <int K, bool WIDE> struct TransPack {
    static constexpr bool kWide = WIDE || K > 4;
    using Row = typename std::conditional<kWide, unsigned long long, unsigned int>::type;
    static constexpr int kBits = (kWide ? 64 : 32) / K;
    static constexpr long long kMaxT = (1ll << kBits) - 1;
    static constexpr bool kFlush = K > 4;
    Row row[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < K; ++i) row[i] = 0;
    }
    // one transition x -> xn; inc = Row(1) << (kBits * xn) is prepared when xn is drawn
    __device__ __forceinline__ void add(int x, Row inc) {
#pragma unroll
        for (int i = 0; i < K; ++i) if (x == i) row[i] += inc;
    }
    __device__ __forceinline__ int get(int i, int j) const { return (int)((row[i] >> (kBits * j)) & (Row)kMaxT); }
};

// What the backward step needs once X_{t+1} = x is known, fetched with one (fp32, K<=3) vector load from shared memory
// instead of K+1 register selects: column x of A and the packed-counter increment of destination x.
template <typename R, int K, bool WIDE> struct alignas(16) GibbsEntry {
    R a[K];
    typename TransPack<K, WIDE>::Row inc;
};
template <typename R, int K, bool WIDE> __host__ __device__ constexpr size_t gibbs_smem_bytes(bool smooth, int n_h);

struct NoSig {};
template <typename R> struct SigChain {
    const R* s0;                  // row j -> s0[j*sld] (z-scale, sign = signal flag)
    long long sld;
    R* opi;                       // pi_end field of the draw being saved (stride ocs between states)
    size_t ocs;
    int pi_back;
};
template <typename R, int K> struct SigBack {
    R Sm[K - 1], Qm[K - 1];       // statistics of the signals of states 0..K-2
    int Mi[K - 1];
    int nback;                    // smoothing steps left before pib[N - pi_back, :] is emitted
};

template <typename R, int K, bool SMOOTH, bool LOGLIK, bool WIDE, bool SIG = false>
struct GibbsWarp {
    static_assert(!(SIG && SMOOTH), "the signals tier has no smoothed-mean variant");
    using Pack = TransPack<K, WIDE>;
    using Row = typename Pack::Row;
    static constexpr int kRing = gibbs_ring_stages<R, K>();          // shadows the global default inside this class

    using Entry = GibbsEntry<R, K, WIDE>;

    // per-sweep constants of one chain
    // (the struct is passed BY VALUE to the out-of-line passes and sits exactly at the 128 bytes (fp32, K = 3) ptxas still passes in
    //  registers: 4 more bytes pushed the kernel's frame from 424 to 2208 bytes and cost 3 % on C2 x 256.  `ysm` therefore lives in the
    //  4-byte alignment hole in front of `yld` — K*K + K + 5 words precede it, always an odd number)
    struct Chain : std::conditional<SIG, SigChain<R>, NoSig>::type {
        R A[K][K];
        R c;                      // shift of the sufficient statistics
        int rank[K];              // position of each chain label in increasing-mu order
        int T, Tw, off;           // window length, warp length, right-alignment offset (row j <-> t = j - off)
        int rag_rows;             // rows [0, rag_rows) of the warp's frame lie before the window of some lane (shorter window or padding lane): steps there are guarded per lane; 0 = all windows equal
        unsigned ysm;             // row 0 of the lane's frame in the shared-memory copy of the series (byte address; y_sm_elems > 0)
        long long yld;
        const R* y0;              // row j -> y0[j*yld]
        R* pi0;                   // row j, state k -> pi0[((j + pad) >> 2)*4*K*32 + k*128 + ((j + pad) & 3)]  (lane*4 folded in)
        R* pacc0;
        R* facc0;                 // in-sample forecast sums (SMOOTH): row j, horizon h -> facc0[(j*n_hi + h)*32]
        unsigned mh_off;          // shared memory (bytes): this thread's A^h mu vectors, mh[(h*K + s)*kGibbsThreads]
        int n_hi;                 // horizons with in-sample forecasts (0 = none)
        unsigned tab_s;           // shared memory, 32-bit shared-window address (ld.shared operand): tab[x*kGibbsThreads] = {A[:,x], 1 << (kBits*x)} of this thread's chain
        unsigned ring_off;        // shared memory (bytes): this warp's ring of kRing groups x 4 rows x K x 32 lanes (HMC_ASYNC)
    };

    // state carried by the backward pass; the sampled state of the later time step is kept one-hot in (lt[])
    struct Back : std::conditional<SIG, SigBack<R, K>, NoSig>::type {
        R Sd[K - 1], Qd[K - 1];   // statistics of states 0..K-2 (the last one follows from the window totals)
        Pack tr;
        Row inc;                  // Row(1) << (kBits * x_{t+1})
        R Acol[K];                // A[:, x_{t+1}]
        R gate;                   // pif[t+1, x_{t+1}]  (quirk Q5)
        R pb[K];                  // smoothed marginal of the later time step (SMOOTH)
    };

    // lt[i] = (cumulative_i < u*total), monotone; state == i  <=>  lt[i-1] && !lt[i]
    static __device__ __forceinline__ bool is_state(const bool (&lt)[K - 1], int i) {
        if (i == 0) return !lt[0];
        if (i == K - 1) return lt[K - 2];
        return lt[i - 1] && !lt[i];
    }
    template <typename V> static __device__ __forceinline__ V pick(const bool (&lt)[K - 1], const V (&v)[K]) {
        V r = v[0];
#pragma unroll
        for (int i = 1; i < K; ++i) r = lt[i - 1] ? v[i] : r;
        return r;
    }
    // Distributions 0.21 rand(Categorical(p)) on unnormalised p: first i with cumulative_i >= u*total
    static __device__ __forceinline__ void draw(const R (&p)[K], R u, bool (&lt)[K - 1]) {
        R cum[K];
        cum[0] = p[0];
#pragma unroll
        for (int i = 1; i < K; ++i) cum[i] = cum[i - 1] + p[i];
        const R thr = u * cum[K - 1];
#pragma unroll
        for (int i = 0; i < K - 1; ++i) lt[i] = cum[i] < thr;
    }

    // book-keeping after X_t was drawn (one-hot in lt): transition x_t -> x_{t+1}, statistics, and the selections
    // the next (earlier) step needs
    // (ALLSIG: every time step of the batch is a signal — the observation statistics are not kept at all)
    template <bool ALLSIG = false>
    static __device__ __forceinline__ void commit(Back& b, const Chain& ch, const bool (&lt)[K - 1], const R (&pt)[K], R yt, R sw, bool first) {
        if (!first) {
#pragma unroll
            for (int i = 0; i < K; ++i) if (is_state(lt, i)) b.tr.row[i] += b.inc;
        }
        const R d = yt - ch.c, dd = d * d;
        bool obs = true;
        if constexpr (SIG) {                                       // signals keep their own statistics (:267-300)
            obs = ALLSIG ? false : !(sw < R(0));
#pragma unroll
            for (int i = 0; i < K - 1; ++i)
                if (!obs && is_state(lt, i)) { b.Sm[i] += d; b.Qm[i] += dd; b.Mi[i] += 1; }
        }
        if constexpr (!(SIG && ALLSIG)) {
#pragma unroll
            for (int i = 0; i < K - 1; ++i)
                if (obs && is_state(lt, i)) { b.Sd[i] += d; b.Qd[i] += dd; }
        }
        select_later(b, ch, lt, pt);
    }
    // what the next (earlier) step needs once this step's state (one-hot in lt) is known: column x of A, the packed-counter
    // increment of destination x, and pif[t, x] (the quirk-Q5 gate)
    static __device__ __forceinline__ void select_later(Back& b, const Chain& ch, const bool (&lt)[K - 1], const R (&pt)[K]) {
#if HMC_SEL_LDS
        // the entry's 32-bit shared-window address is selected among K per-thread constants (K-1 selects) and read with
        // ld.shared: stepping a generic pointer cost 5 integer instructions per step (offset chain + window base add)
        unsigned ea = ch.tab_s;
#pragma unroll
        for (int i = 1; i < K; ++i) ea = lt[i - 1] ? ch.tab_s + (unsigned)(i * kGibbsThreads * sizeof(Entry)) : ea;
        Entry en;
        lds_struct(ea, en);
        b.inc = en.inc;
#pragma unroll
        for (int r = 0; r < K; ++r) b.Acol[r] = en.a[r];
#else
        Row incs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) incs[i] = (Row)1 << (Pack::kBits * i);
        b.inc = pick<Row>(lt, incs);
#pragma unroll
        for (int r = 0; r < K; ++r) b.Acol[r] = pick<R>(lt, ch.A[r]);
#endif
        b.gate = pick<R>(lt, pt);
    }

    // one backward step at loop row j: draw X_t | X_{t+1} (src/Hmc.jl:466-481 in the pif form), update the statistics.
    // Quirk Q5 (:472-480: p = 1/D when total = pif[t+1, x_{t+1}] <= eps()) practically never fires.  GATED = handle it
    // in place; otherwise only record it in `bad` and let the caller redo the pass gated (counter-based RNG: same draws).
    // in-sample forecast of the row whose smoothed marginal is pb: sum_s pb[s] (A^h mu)[s] for every horizon (:683-699)
    static __device__ __forceinline__ void insample_accumulate(const Chain& ch, const R (&pb)[K], R* __restrict__ fap) {
        const R* mh = reinterpret_cast<const R*>(smem_base() + ch.mh_off);
        for (int j = 0; j < ch.n_hi; ++j) {
            R f = R(0);
#pragma unroll
            for (int s = 0; s < K; ++s) f = fma(pb[s], mh[(j * K + s) * kGibbsThreads], f);
            st_stream(fap + j * 32, ld_stream(fap + j * 32) + f);
        }
    }

    // NBACK (signals tier): this step may still belong to the pi_back smoothing steps at the end of the window — compiled out of
    // the hot loops, which only start once those steps are behind every lane of the warp
    template <bool GATED, bool NBACK = true, bool ALLSIG = false>
    static __device__ __forceinline__ void back_step(Back& b, const Chain& ch, const R (&pt)[K], R* __restrict__ pap, R* __restrict__ fap,
                                                      R yt, R sw, uint32_t word, bool save, bool& bad) {
        R p[K];
#pragma unroll
        for (int r = 0; r < K; ++r) p[r] = pt[r] * b.Acol[r];
        if (GATED) {
            if (!(b.gate > Real<R>::eps())) {
#pragma unroll
                for (int r = 0; r < K; ++r) p[r] = R(1);
            }
        } else {
            bad = bad || !(b.gate > Real<R>::eps());
        }
        bool lt[K - 1];
        draw(p, u01<R>(word), lt);
        if (SMOOTH) {
            smooth_step<R, K>(ch.A, pt, b.pb);
            if (save) {
#pragma unroll
                for (int s = 0; s < K; ++s) st_stream(pap + ch.rank[s] * 32, ld_stream(pap + ch.rank[s] * 32) + b.pb[s]);
                insample_accumulate(ch, b.pb, fap);
            }
        }
        if constexpr (SIG && NBACK) {
            // samples.πb[:, endIndex, :] (:893): the smoothed marginal pi_back rows before the end of the window
            if (b.nback > 0) {
                smooth_step<R, K>(ch.A, pt, b.pb);
                if (--b.nback == 0 && save) {
#pragma unroll
                    for (int s = 0; s < K; ++s) ch.opi[(size_t)ch.rank[s] * ch.ocs] = b.pb[s];
                }
            }
        }
        commit<ALLSIG>(b, ch, lt, pt, yt, sw, false);
    }

    // forward filter (forwardupdate_P! :371-440); pif rows stored for the backward pass.  CHECKED = per-step handling
    // of a zero / non-finite normaliser (the reference only warns, :435); the hot path runs unchecked and is re-run
    // checked when the final row is not finite (a NaN, once produced, propagates to the last row).
    // The two passes are separate out-of-line functions (arguments and results by value): each gets its own register
    // allocation, and the per-sweep state of the caller is saved around the call instead of squeezing the hot loops.
    struct Vec { R v[K]; };
    struct FwdOut { Vec pf; R ll; int events; };
    // FACC (HMCGPU_FLAG_FILTERED_MEAN, a saved draw): every filtered row is also added to the per-date sums (sorted labels) and
    // its forecasts pif[t,:]' A^h mu to the in-sample forecast sums — the table the reference published (forecats_insample.csv).
    // (YSM: the observations come from the shared-memory copy of the series — hot single-series variants only)
    template <bool RAGGED, bool CHECKED, bool STREAM = false, bool FACC = false, bool YSM = false>
    static __device__ HMC_FWD_ATTR FwdOut forward_pass(const Chain ch, const Emission<R, K> em, const Vec rho_in) {
        static_assert(!YSM || (!STREAM && !CHECKED), "shared-memory observations: unit stride");
        FwdOut o;
        R (&pf)[K] = o.pf.v;
        R& ll = o.ll;
        const R (&rho)[K] = rho_in.v;
        const long long yld = (!STREAM && !CHECKED) ? 1ll : ch.yld;   // the hot non-streaming variants only run when every window shares one series (stride 1): immediate offsets instead of address arithmetic
        constexpr bool ragged = RAGGED;
        int events = 0;
#pragma unroll
        for (int s = 0; s < K; ++s) pf[s] = rho[s];                  // t = 1 uses ρ (:390)
        ll = R(0);
        YRef<R, YSM> yp;
        if constexpr (YSM) yp.a = ch.ysm; else yp.p = ch.y0;
        const R* sp = nullptr;                                       // z-scale of the emission per row (SIG)
        if constexpr (SIG) sp = ch.s0;
        R* pip = ch.pi0;
        // fp32: the K states are processed as K/2 packed pairs (+ one scalar state when K is odd) with FFMA2/FMUL2/FADD2:
        // the kernel is bound by instruction issue, and the packed forms halve the slots of the emission quadratic, the
        // K x K prediction and the normalisation.  Same operations and roundings as the scalar form (Emission/forward_step).
        constexpr int KP = K / 2;
        constexpr bool kOdd = (K & 1) != 0;
        f2 negmu2[KP > 0 ? KP : 1], q2[KP > 0 ? KP : 1], c2[KP > 0 ? KP : 1], A2[K][KP > 0 ? KP : 1];
        float negmu_l = 0.f, q_l = 0.f, c_l = 0.f, A_l[K];
        if constexpr (sizeof(R) == 4) {
#pragma unroll
            for (int p = 0; p < KP; ++p) {
                negmu2[p] = mk2(-(float)em.mu[2 * p], -(float)em.mu[2 * p + 1]);
                q2[p] = mk2((float)em.q[2 * p], (float)em.q[2 * p + 1]);
                c2[p] = mk2((float)em.c[2 * p], (float)em.c[2 * p + 1]);
#pragma unroll
                for (int r = 0; r < K; ++r) A2[r][p] = mk2((float)ch.A[r][2 * p], (float)ch.A[r][2 * p + 1]);
            }
            if (kOdd) {
                negmu_l = -(float)em.mu[K - 1]; q_l = (float)em.q[K - 1]; c_l = (float)em.c[K - 1];
#pragma unroll
                for (int r = 0; r < K; ++r) A_l[r] = (float)ch.A[r][K - 1];
            }
        }
        R buf[4][K];                                                 // the 4 rows of the current tile (stored together)
        // one step at row j = position u of its tile; yo / so: offsets (in rows) of y and of the z-scale from yp / sp
        // (guard: a std::integral_constant — the per-lane window test is compiled out for the rows where every lane is inside its window)
        auto step = [&](auto guard, int j, int u, int yo, R ypre, bool preloaded, R swpre = R(1)) {
            if (!decltype(guard)::value || j >= ch.off) {
                const R yt = preloaded ? ypre : (yp + yo * yld).ld();   // STREAM: loaded one iteration ahead by the caller
                R sw = R(1);                                         // signals: sd x (1+kappa) (:382) <=> z scaled by 1/(1+kappa)
                if constexpr (SIG) sw = preloaded ? swpre : ld_ro(sp + yo * ch.sld);   // STREAM: fetched with the observation
                if constexpr (sizeof(R) == 4) {
                    const f2 y2 = splat2((float)yt);
                    const f2 sw2 = splat2((float)sw);
                    f2 l2[KP > 0 ? KP : 1];
                    float l_l = -3.0e38f;
#pragma unroll
                    for (int p = 0; p < KP; ++p) {
                        f2 d = y2 + negmu2[p];
                        if constexpr (SIG) d = d * sw2;
                        l2[p] = fma2(d * d, q2[p], c2[p]);
                    }
                    if (kOdd) {
                        float d = (float)yt + negmu_l;
                        if constexpr (SIG) d *= (float)sw;
                        l_l = fmaf(d * d, q_l, c_l);
                    }
                    float m2 = l_l;
#pragma unroll
                    for (int p = 0; p < KP; ++p) m2 = fmaxf(m2, fmaxf(l2[p].v.x, l2[p].v.y));
                    const f2 nm = splat2(-m2);
                    f2 qq2[KP > 0 ? KP : 1];
                    float qq_l = 0.f;
#pragma unroll
                    for (int p = 0; p < KP; ++p) {
                        const f2 a = l2[p] + nm;
                        const f2 e = mk2(Real<float>::ex2(a.v.x), Real<float>::ex2(a.v.y));
                        f2 pred = splat2((float)pf[0]) * A2[0][p];
#pragma unroll
                        for (int r = 1; r < K; ++r) pred = fma2(splat2((float)pf[r]), A2[r][p], pred);
                        qq2[p] = pred * e;
                    }
                    if (kOdd) {
                        const float e = Real<float>::ex2(l_l - m2);
                        float pred = (float)pf[0] * A_l[0];
#pragma unroll
                        for (int r = 1; r < K; ++r) pred = fmaf((float)pf[r], A_l[r], pred);
                        qq_l = pred * e;
                    }
                    float tot = kOdd ? qq_l : 0.f;
#pragma unroll
                    for (int p = 0; p < KP; ++p) tot += qq2[p].v.x + qq2[p].v.y;
                    const f2 inv = splat2(Real<float>::rcp(tot));
#pragma unroll
                    for (int p = 0; p < KP; ++p) { const f2 r2 = qq2[p] * inv; pf[2 * p] = (R)r2.v.x; pf[2 * p + 1] = (R)r2.v.y; }
                    if (kOdd) pf[K - 1] = (R)(qq_l * inv.v.x);
                    if (CHECKED && !((tot > 0.f) && (tot < 3.0e38f))) {
                        ++events;
#pragma unroll
                        for (int s = 0; s < K; ++s) pf[s] = R(1) / R(K);
                    }
                    if (LOGLIK) {
                        float l2t = Real<float>::lg2(tot) + m2;
                        if constexpr (SIG) l2t += Real<float>::lg2(fabsf((float)sw));   // the 1/(1+kappa) of the signal pdf
                        ll += (R)(l2t * 0.6931471805599453f);
                    }
                } else {
                    R e[K];
                    if constexpr (SIG) em.eval_scaled(yt, sw, e); else em.eval(yt, e);
                    bool ok;
                    const R tot = forward_step<R, K>(ch.A, e, pf, ok);
                    if (CHECKED && !ok) {
                        ++events;
#pragma unroll
                        for (int s = 0; s < K; ++s) pf[s] = R(1) / R(K);
                    }
                    if (LOGLIK) ll += (R)log((double)tot);
                }
#pragma unroll
                for (int s = 0; s < K; ++s) buf[u][s] = pf[s];
                if constexpr (FACC) {
                    R* pap = ch.pacc0 + (size_t)j * K * 32;
#pragma unroll
                    for (int s = 0; s < K; ++s) st_stream(pap + ch.rank[s] * 32, ld_stream(pap + ch.rank[s] * 32) + pf[s]);
                    insample_accumulate(ch, pf, ch.facc0 + (size_t)j * ch.n_hi * 32);
                }
            }
        };
        auto store_tile = [&]() {
#pragma unroll
            for (int s = 0; s < K; ++s) st_quad(pip + s * 128, buf[0][s], buf[1][s], buf[2][s], buf[3][s]);
        };
        const int pad = (4 - (ch.Tw & 3)) & 3;
        int j = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int s = 0; s < K; ++s) buf[u][s] = R(0);
        if (pad != 0) {   // tile 0: its first `pad` positions are padding (rows are right-aligned to a multiple of 4); a frame of whole tiles starts in the loops below
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = u - pad;
                if (jj >= 0 && jj < ch.Tw) step(std::integral_constant<bool, RAGGED>{}, jj, u, jj, R(0), false);
            }
            store_tile();
            j = (4 - pad < ch.Tw) ? 4 - pad : ch.Tw;
            yp += (long long)j * yld; pip += 4 * K * 32;
            if constexpr (SIG) sp += (long long)j * ch.sld;
        }
        // STREAM: the observations of the next 4 steps are fetched into registers one iteration ahead (from lines that
        // were pulled into L1 kYAhead steps ahead), so neither the DRAM nor the L1 latency sits in the dependent chain
        R yn[4] = {R(0), R(0), R(0), R(0)}, sn[4] = {R(1), R(1), R(1), R(1)};
        auto load4 = [&](int jj, const YRef<R, YSM> p, const R* q) {  // observations (and, SIG, their z-scales: one more streamed array per series)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = jj + u < ch.Tw && (!ragged || jj + u >= ch.off);
                yn[u] = in ? (p + u * yld).ld() : R(0);
                if constexpr (SIG) sn[u] = in ? ld_ro(q + u * ch.sld) : R(1);
            }
        };
        if constexpr (STREAM) load4(j, yp, sp);
        // the tiles whose rows lie before the window of some lane run the guarded step, the rest of the frame the plain one
        auto run_tiles = [&](auto guard, const int until) {
            for (; j + 3 < ch.Tw && j < until; j += 4, yp += 4 * yld, pip += 4 * K * 32) {
                if constexpr (STREAM) {
                    if (j + kYAhead + 3 < ch.Tw && (!ragged || j + kYAhead >= ch.off)) {   // rows of this lane's own window only
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            prefetch_l1((yp + (kYAhead + u) * yld).gptr());
                            if constexpr (SIG) { if (ch.sld != 1) prefetch_l1(sp + (kYAhead + u) * ch.sld); }   // (a mask shared by every series stays in L1)
                        }
                    }
                    const R c0 = yn[0], c1 = yn[1], c2v = yn[2], c3 = yn[3], z0 = sn[0], z1 = sn[1], z2 = sn[2], z3 = sn[3];
                    const R* spn = sp;
                    if constexpr (SIG) spn = sp + 4 * ch.sld;
                    load4(j + 4, yp + 4 * yld, spn);
                    step(guard, j, 0, 0, c0, true, z0); step(guard, j + 1, 1, 1, c1, true, z1); step(guard, j + 2, 2, 2, c2v, true, z2); step(guard, j + 3, 3, 3, c3, true, z3);
                } else {
                    step(guard, j, 0, 0, R(0), false); step(guard, j + 1, 1, 1, R(0), false); step(guard, j + 2, 2, 2, R(0), false); step(guard, j + 3, 3, 3, R(0), false);
                }
                store_tile();
                if constexpr (SIG) sp += 4 * ch.sld;
            }
        };
        if constexpr (RAGGED) run_tiles(std::true_type{}, ch.rag_rows);
        run_tiles(std::false_type{}, ch.Tw);                        // (no tail: the rows end on a tile boundary)
        o.events = events;
        return o;
    }
    // backward state sampling (update_X! :459-484) fused with the next sweep's statistics (update_μσ! :254-258/:291-294,
    // update_A! :362-365) and, optionally, backwardupdate_P! (:442-457).
    // The i-th uniform consumed (i = 0 for X[N]) is word i&3 of Philox block i>>2 of this sweep.
    struct NoAcc {};
    struct TransAcc { int n[K * K]; };                               // flushed transition counts (K > 4)
    struct BackOut : std::conditional<Pack::kFlush, TransAcc, NoAcc>::type { Back b; int xN; bool bad; bool lost; };   // lost: a bulk copy never completed (HMC_TMA)
    template <bool RAGGED, bool GATED, bool STREAM = false, bool ALLSIG = false, bool YSM = false>
    static __device__ HMC_BACK_ATTR BackOut backward_pass(const Chain ch, const Vec pf_in, const RngKey key, const uint32_t sweep,
                                                         const unsigned flags, const bool save) {
        BackOut o;
        Back b;                      // a local (registers): `o` is returned through memory when K > 4
        const R (&pf)[K] = pf_in.v;
        bool bad = false, lost = false;
        int xN = 0;
        const int Tw = ch.Tw, T = ch.T;
        const long long ys = (!STREAM && !GATED) ? 1ll : ch.yld;       // see forward_pass
        constexpr bool ragged = RAGGED;
#pragma unroll
        for (int i = 0; i < K - 1; ++i) { b.Sd[i] = R(0); b.Qd[i] = R(0); }
        b.tr.clear();
        int since = 8;                                               // steps since the packed counters were last flushed (+ slack)
        // The flushed counts are only touched every kMaxT steps: the rolled inner loops index them dynamically, which keeps
        // them in local memory instead of 64 registers that would otherwise live through the hot loop.
        if constexpr (Pack::kFlush) {
#pragma unroll 1
            for (int q = 0; q < K * K; ++q) o.n[q] = 0;
        }
        auto flush = [&]() {
            if constexpr (Pack::kFlush) {
                Row tmp[K];
#pragma unroll
                for (int r = 0; r < K; ++r) tmp[r] = b.tr.row[r];
#pragma unroll 1
                for (int r = 0; r < K; ++r) {                        // rolled over rows (dynamic index -> local memory) ...
                    const Row v = tmp[r];
#pragma unroll
                    for (int c2 = 0; c2 < K; ++c2)                   // ... K independent read-modify-writes per row
                        o.n[r * K + c2] += (int)((v >> (Pack::kBits * c2)) & (Row)Pack::kMaxT);
                }
                b.tr.clear();
                since = 8;
            }
        };
        b.inc = 0; b.gate = R(1);
#pragma unroll
        for (int s = 0; s < K; ++s) { b.Acol[s] = R(0); b.pb[s] = R(0); }
        static_assert(!YSM || (!STREAM && !GATED), "shared-memory observations: unit stride");
        YRef<R, YSM> yp;
        if constexpr (YSM) yp.a = ch.ysm + (unsigned)((Tw - 1) * (int)sizeof(R)); else yp.p = ch.y0 + (long long)(Tw - 1) * ys;
        const R* sp = nullptr;
        long long ss = 0;                                             // stride of the z-scale rows (SIG)
        if constexpr (SIG) {
            ss = ch.sld;
            sp = ch.s0 + (long long)(Tw - 1) * ss;
#pragma unroll
            for (int i = 0; i < K - 1; ++i) { b.Sm[i] = R(0); b.Qm[i] = R(0); b.Mi[i] = 0; }
            b.nback = ch.pi_back < T - 1 ? ch.pi_back : (T > 0 ? T - 1 : 0);
        }
        const int pad = (4 - (Tw & 3)) & 3;                          // tiles of 4 rows, right-aligned (see st_quad)
        auto row_ptr = [&](int jrow) -> const R* {                   // state 0 of row jrow (state s: + s*128)
            const int jp = jrow + pad;
            return ch.pi0 + (size_t)(jp >> 2) * (4 * K * 32) + (jp & 3);
        };
        R* pap = SMOOTH ? ch.pacc0 + (size_t)(Tw - 1) * K * 32 : nullptr;
        uint4 w = rng_block_states(key, sweep, 0u);
        if (T > 0) {
            // X[N] ~ Categorical(pif[N,:]); with quirk Q1 the relabelled row is used with chain labels (:512-514)
            R pN[K];
            if (flags & 1u) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    R vsel = R(0);
#pragma unroll
                    for (int s = 0; s < K; ++s) vsel = (ch.rank[s] == k) ? pf[s] : vsel;
                    pN[k] = vsel;
                }
            } else {
#pragma unroll
                for (int s = 0; s < K; ++s) pN[s] = pf[s];
            }
            bool lt[K - 1];
            draw(pN, u01<R>(w.x), lt);
#pragma unroll
            for (int i = 0; i < K - 1; ++i) xN += lt[i] ? 1 : 0;
            if (SMOOTH || SIG) {
#pragma unroll
                for (int s = 0; s < K; ++s) b.pb[s] = pf[s];           // pib[N,:] = pif[N,:]
            }
            if (SMOOTH) {
                if (save) {
#pragma unroll
                    for (int s = 0; s < K; ++s) st_stream(pap + ch.rank[s] * 32, ld_stream(pap + ch.rank[s] * 32) + pf[s]);
                    insample_accumulate(ch, b.pb, ch.facc0 + (size_t)(Tw - 1) * ch.n_hi * 32);
                }
            }
            R swN = R(1);
            if constexpr (SIG) swN = ld_ro(sp);
            commit<ALLSIG>(b, ch, lt, pf, yp.ld(), swN, true);
        }
        int i = 1;
        // one step; u = position inside the current group of 4 (pointers move once per group).  The filtered rows and
        // observations of a group are loaded into registers one whole group ahead (they do not depend on the sampled
        // states), so the HBM/L2 latency of group g+1 hides behind the dependent chain of group g.
        auto load_sw = [&](int u) -> R {
            if constexpr (SIG) return (!ragged || i + u < T) ? ld_ro(sp - (u + 1) * ss) : R(1);
            else return R(1);
        };
        auto load_row = [&](int u, R (&pt)[K], R& yt, R& st) {
#pragma unroll
            {
                const R* rp = row_ptr(Tw - 1 - (i + u));
#pragma unroll
                for (int s = 0; s < K; ++s) pt[s] = ld_stream(rp + s * 128);
            }
            yt = (!ragged || i + u < T) ? (yp - (u + 1) * ys).ld() : R(0);
            st = load_sw(u);
        };
#define HMC_BACK_NB(NB, u, word, PT, YT, ST)                                                                           \
    if (!ragged || i + (u) < T)                                                                                          \
        back_step<GATED, NB, ALLSIG>(b, ch, PT, SMOOTH ? pap - ((u) + 1) * K * 32 : nullptr,                             \
                                     SMOOTH ? ch.facc0 + (size_t)(Tw - 1 - (i + (u))) * ch.n_hi * 32 : nullptr, YT, ST, (word), save, bad);
#define HMC_BACK(u, word, PT, YT, ST) HMC_BACK_NB(true, u, word, PT, YT, ST)
        {
            R p0[K], p1[K], p2[K], y0, y1, y2, s0 = R(1), s1 = R(1), s2 = R(1);
            if (Tw > 1) load_row(0, p0, y0, s0);
            if (Tw > 2) load_row(1, p1, y1, s1);
            if (Tw > 3) load_row(2, p2, y2, s2);
            if (Tw > 1) { HMC_BACK(0, w.y, p0, y0, s0) }
            if (Tw > 2) { HMC_BACK(1, w.z, p1, y1, s1) }
            if (Tw > 3) { HMC_BACK(2, w.w, p2, y2, s2) }
            const int done = Tw > 3 ? 3 : Tw - 1;
            i += done; yp -= done * ys; if (SMOOTH) pap -= (size_t)done * K * 32;
            if constexpr (SIG) sp -= done * ss;
        }
        if constexpr (SIG) {
            // the remaining pi_back smoothing steps (:893), whole groups of 4 straight from global memory: the ring loop below
            // then runs without the smoother (warp-uniform: the pass is entered by the whole warp)
            while (i + 3 < Tw && __any_sync(0xffffffffu, b.nback > 0)) {
                R c0[K], c1[K], c2[K], c3[K], y0, y1, y2, y3, s0 = R(1), s1 = R(1), s2 = R(1), s3 = R(1);
                load_row(0, c0, y0, s0); load_row(1, c1, y1, s1); load_row(2, c2, y2, s2); load_row(3, c3, y3, s3);
                w = rng_block_states(key, sweep, (uint32_t)(i >> 2));
                HMC_BACK(0, w.x, c0, y0, s0) HMC_BACK(1, w.y, c1, y1, s1) HMC_BACK(2, w.z, c2, y2, s2) HMC_BACK(3, w.w, c3, y3, s3)
                i += 4; yp -= 4 * ys; sp -= 4 * ss;
            }
        }
#if HMC_ASYNC
        {
            // Groups of 4 rows are contiguous in the warp's tile (4*K*32 elements).  They are copied global -> shared
            // kRing-1 groups ahead with 16-byte cp.async (every lane moves K*sizeof(R)/4 chunks per group), so the HBM/L2
            // latency never sits in the dependent chain of the sampler; the rows are then read back with conflict-free LDS.
            constexpr int kGroupElems = 4 * K * 32;
            constexpr int kChunksPerLane = (int)(kGroupElems * sizeof(R) / 16 / 32);
            const int lane = threadIdx.x & 31;
            R* const ring = reinterpret_cast<R*>(smem_base() + ch.ring_off);
            const int n_groups = (Tw - i) / 4;                          // full groups (i == 4 here when there are any)
            // group g (0-based) holds rows [jlo, jlo+3], jlo = Tw - 8 - 4g  (the rows of steps i = 4+4g .. 7+4g, highest first)
            // running source pointer / ring stage of the next group to fetch (loop-carried: re-deriving them from g cost ~20
            // integer instructions per group)
            constexpr bool kTma = HMC_TMA && K <= 4;
            const char* nsrc = reinterpret_cast<const char*>(ch.pi0 - lane * 4 + (long long)(Tw + pad - 4 - i) * K * 32) + (kTma ? 0 : lane * 16);   // tile of group 0 (rows of steps i..i+3; i is a multiple of 4 here)
            const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + (kTma ? 0u : (unsigned)lane * 16u);
            constexpr unsigned kGroupBytes = (unsigned)(kGroupElems * sizeof(R));
            const unsigned ring_rd = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)(lane * 4 * sizeof(R));   // this lane's rows of state 0, stage 0
            unsigned nstage = 0;                                         // byte offset of the stage the next fetch fills
            int nleft = n_groups;                                        // groups not fetched yet
            // TMA: one mbarrier per ring stage of this warp, behind the rings; (re)initialised per pass.  The forward pass wrote the
            // tiles through the generic proxy: a proxy fence orders those writes before the bulk copies read them.  Everything the
            // copy is issued from is broadcast from lane 0 first: the compiler then keeps source, destination, barrier and counters
            // in uniform registers and emits ONE UBLKCP per group (per-lane operands cost an ELECT/R2UR waterfall loop per copy).
            unsigned long long usrc = 0;                                 // (uniform) source of the next fetch
            unsigned uring = 0, ubar0 = 0, ufetch = 0;                   // (uniform) ring base, first mbarrier, groups fetched so far
            int uleft = 0;
            if constexpr (kTma) {
                const unsigned wq = threadIdx.x >> 5;
                const unsigned bar_l = (unsigned)__cvta_generic_to_shared(smem_base()) + ch.ring_off
                                       + (unsigned)(sizeof(R) * ((kGibbsThreads / 32) - wq) * (kRing * 4 * K * 32)) + wq * (kRing * 8u);
                const unsigned long long src_l = (unsigned long long)nsrc;
                usrc = ((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(src_l >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (unsigned)src_l, 0);
                uring = __shfl_sync(0xffffffffu, ring_s, 0);
                ubar0 = __shfl_sync(0xffffffffu, bar_l, 0);
                uleft = __shfl_sync(0xffffffffu, n_groups, 0);
                fence_proxy_async();
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < kRing; ++q) mbar_init(ubar0 + 8u * q, 1u);
                    fence_mbar_init();
                }
                __syncwarp();
            }
            // HMC_TMA = groups per bulk copy G (1 or 2): the ring holds kRing/G copy stages, each with its own mbarrier.  The groups
            // of one copy are contiguous in memory in DESCENDING step order, so inside a stage of G = 2 the later group comes first.
            constexpr int kG = kTma ? HMC_TMA : 1, kCopies = kRing / kG;
            static_assert(!kTma || (kRing % kG == 0 && kCopies >= 2), "ring stages must hold whole copies");
            auto issue = [&]() {                                         // TMA: one call per COPY (G groups); cp.async: per group
                if constexpr (kTma) {
                    if (uleft > 0) {
                        const unsigned n = uleft >= kG ? (unsigned)kG : 1u;          // groups in this copy (the last one may be short)
                        if (elect_one()) {
                            mbar_expect_tx(ubar0 + 8u * ufetch, kGroupBytes * n);
                            bulk_g2s(uring + kGroupBytes * kG * ufetch + kGroupBytes * ((unsigned)kG - n),
                                     reinterpret_cast<const void*>(usrc - kGroupBytes * (n - 1u)), kGroupBytes * n, ubar0 + 8u * ufetch);
                        }
                    }
                    ufetch = (ufetch + 1u == (unsigned)kCopies) ? 0u : ufetch + 1u;
                    uleft -= kG;
                    usrc -= kGroupBytes * kG;
                } else {
                    if (nleft > 0) {
#pragma unroll
                        for (int m = 0; m < kChunksPerLane; ++m) cp_async16_s(ring_s + nstage + 512u * m, nsrc + 512 * m);
                    }
                    cp_async_commit();                                   // (possibly empty) keeps the group count uniform
                    --nleft;
                    nsrc -= kGroupBytes;
                    nstage = (nstage + kGroupBytes == kRing * kGroupBytes) ? 0u : nstage + kGroupBytes;
                }
            };
            unsigned uread = 0, uparity = 0;                             // (uniform, TMA) copy stage being consumed and its phase parity
#pragma unroll
            for (int g = 0; g < (kTma ? kCopies : kRing) - 1; ++g) issue();
            unsigned rstage = 0;                                         // byte offset of the stage read in this iteration
            R ynx[4] = {R(0), R(0), R(0), R(0)}, snx[4] = {R(1), R(1), R(1), R(1)};
            auto loady4 = [&](int ii, const YRef<R, YSM> p, const R* q) {   // observations (SIG: and z-scales) of steps ii..ii+3 (rows below p / q)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ynx[u] = (!ragged || ii + u < T) ? (p - (u + 1) * ys).ld() : R(0);
                    if constexpr (SIG) snx[u] = (!ragged || ii + u < T) ? ld_ro(q - (u + 1) * ss) : R(1);
                }
            };
            if constexpr (STREAM) { if (n_groups > 0) loady4(i, yp, sp); }
            for (int g = 0; g < n_groups; ++g, i += 4, yp -= 4 * ys, pap -= SMOOTH ? 4 * K * 32 : 0, sp -= 4 * ss) {
                if constexpr (Pack::kFlush) { if ((since += 4) > Pack::kMaxT) flush(); }
                if (STREAM && i + kYAhead + 3 < T) {                       // rows of steps i+kYAhead .. i+kYAhead+3 (this lane's window)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        prefetch_l1((yp - (kYAhead + u + 1) * ys).gptr());
                        if constexpr (SIG) { if (ss != 1) prefetch_l1(sp - (kYAhead + u + 1) * ss); }
                    }
                }
                if (!kTma || kG == 1 || (g & (kG - 1)) == 0) issue();    // group g + kRing - 1 (TMA: the copy kCopies-1 ahead): overwrites what iteration g-1 finished with
                if constexpr (kTma) {
                    if (kG == 1 || (g & (kG - 1)) == 0) {                // first group of a copy: wait for it
                        const unsigned bar = ubar0 + 8u * uread;
                        if (!mbar_try_wait(bar, uparity)) lost |= !mbar_wait_bounded(bar, uparity);   // the first poll almost always succeeds (issued kCopies-1 copies ago)
                        if (++uread == (unsigned)kCopies) { uread = 0u; uparity ^= 1u; }
                    }
                } else {
                    cp_async_wait<kRing - 1>();                          // group g has landed (for this lane's chunks)
                    __syncwarp();                                        // ... and for every other lane's
                }
                const unsigned roff = (kG == 2 ? ((g & 1) ? rstage - kGroupBytes : rstage + kGroupBytes) : rstage);
                const R* st = reinterpret_cast<const R*>(reinterpret_cast<const char*>(ring) + roff) + lane * 4;
                rstage = (rstage + kGroupBytes == kRing * kGroupBytes) ? 0u : rstage + kGroupBytes;
                R c0[K], c1[K], c2[K], c3[K], y0, y1, y2, y3, s0 = R(1), s1 = R(1), s2 = R(1), s3 = R(1);
#pragma unroll
                for (int s = 0; s < K; ++s) {                            // position 3 = highest row = first step
                    if constexpr (HMC_RING_LDS32) lds_quad(ring_rd + roff + (unsigned)(s * 128 * sizeof(R)), c3[s], c2[s], c1[s], c0[s]);
                    else ld_quad_shared(st + s * 128, c3[s], c2[s], c1[s], c0[s]);
                }
                if constexpr (STREAM) {                                  // fetched one group ahead (see forward_pass)
                    y0 = ynx[0]; y1 = ynx[1]; y2 = ynx[2]; y3 = ynx[3];
                    s0 = snx[0]; s1 = snx[1]; s2 = snx[2]; s3 = snx[3];
                    if (g + 1 < n_groups) loady4(i + 4, yp - 4 * ys, sp - 4 * ss);
                } else {
                    y0 = (!ragged || i + 0 < T) ? (yp - 1 * ys).ld() : R(0);
                    y1 = (!ragged || i + 1 < T) ? (yp - 2 * ys).ld() : R(0);
                    y2 = (!ragged || i + 2 < T) ? (yp - 3 * ys).ld() : R(0);
                    y3 = (!ragged || i + 3 < T) ? (yp - 4 * ys).ld() : R(0);
                }
                if constexpr (!STREAM) { s0 = load_sw(0); s1 = load_sw(1); s2 = load_sw(2); s3 = load_sw(3); }
                w = rng_block_states(key, sweep, (uint32_t)(i >> 2));
                HMC_BACK_NB(false, 0, w.x, c0, y0, s0) HMC_BACK_NB(false, 1, w.y, c1, y1, s1) HMC_BACK_NB(false, 2, w.z, c2, y2, s2) HMC_BACK_NB(false, 3, w.w, c3, y3, s3)
                __syncwarp();                                            // all lanes are done with this stage
            }
            if constexpr (kTma) {
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < kRing; ++q) mbar_inval(ubar0 + 8u * q);
                }
                __syncwarp();
            } else {
                cp_async_wait<0>();
            }
        }
#else
        for (; i + 3 < Tw; i += 4, yp -= 4 * ys, pap -= SMOOTH ? 4 * K * 32 : 0, sp -= 4 * ss) {
            if constexpr (Pack::kFlush) { if ((since += 4) > Pack::kMaxT) flush(); }
            R c0[K], c1[K], c2[K], c3[K], y0, y1, y2, y3, s0, s1, s2, s3;
            load_row(0, c0, y0, s0); load_row(1, c1, y1, s1); load_row(2, c2, y2, s2); load_row(3, c3, y3, s3);
            w = rng_block_states(key, sweep, (uint32_t)(i >> 2));
            HMC_BACK(0, w.x, c0, y0, s0) HMC_BACK(1, w.y, c1, y1, s1) HMC_BACK(2, w.z, c2, y2, s2) HMC_BACK(3, w.w, c3, y3, s3)
        }
#endif
        if (i < Tw) {
            w = rng_block_states(key, sweep, (uint32_t)(i >> 2));
            R p0[K], p1[K], p2[K], y0 = R(0), y1 = R(0), y2 = R(0), s0 = R(1), s1 = R(1), s2 = R(1);
            load_row(0, p0, y0, s0);
            if (i + 1 < Tw) load_row(1, p1, y1, s1);
            if (i + 2 < Tw) load_row(2, p2, y2, s2);
            HMC_BACK(0, w.x, p0, y0, s0)
            if (i + 1 < Tw) { HMC_BACK(1, w.y, p1, y1, s1) }
            if (i + 2 < Tw) { HMC_BACK(2, w.z, p2, y2, s2) }
        }
#undef HMC_BACK
#undef HMC_BACK_NB
        flush();
        o.b = b;
        o.xN = xN;
        o.bad = bad;
        o.lost = lost;
        return o;
    }

    static __device__ __forceinline__ void run(const GibbsArgs& a, const int warp, const int lane, Entry* smem_tab) {
        const int slot = warp * 32 + lane;
        const int ns = a.n_slots;
        Chain ch;
        ch.T = a.T[slot];
        ch.Tw = a.warp_T[warp];
        ch.off = ch.Tw - ch.T;
        ch.yld = a.yld;
        ch.pi0 = reinterpret_cast<R*>(a.pi) + a.warp_pi_off[warp] + lane * 4;
        const bool filt = !SMOOTH && (a.flags & 64u /*HMCGPU_FLAG_FILTERED_MEAN*/) != 0u;   // filtered means: accumulated by the forward pass
        const bool accum = SMOOTH || filt;
        ch.pacc0 = accum ? reinterpret_cast<R*>(a.pib_acc) + a.warp_pi_off[warp] + lane : nullptr;
        ch.n_hi = (accum && a.fc_acc) ? a.n_h : 0;
        ch.facc0 = ch.n_hi ? reinterpret_cast<R*>(a.fc_acc) + a.warp_pi_off[warp] / K * ch.n_hi + lane : nullptr;
        ch.mh_off = (unsigned)(gibbs_smem_bytes<R, K, WIDE>(false, 0) + threadIdx.x * sizeof(R));
        // realised future observations of this chain (forecast errors), cached once per launch: yf[j*threads + tid] for the
        // j-th sorted horizon — the per-draw global loads sat in front of every forecast store
        const unsigned yf_off = (unsigned)(gibbs_smem_bytes<R, K, WIDE>(false, 0) + (accum ? sizeof(R) * (size_t)a.n_h * K * kGibbsThreads : 0)
                                           + threadIdx.x * sizeof(R));
        {
            R* yf = reinterpret_cast<R*>(smem_base() + yf_off);
            for (int j = 0; j < a.n_h; ++j) yf[j * kGibbsThreads] = reinterpret_cast<const R*>(a.yfut)[(size_t)a.h_slot[j] * ns + slot];
        }
        ch.y0 = reinterpret_cast<const R*>(a.y) + a.ybase[slot] - (long long)ch.off * ch.yld;
        // (the copy of the series sits behind everything else in the block's shared memory; rows in front of a ragged lane's window
        //  are never read, so the address may point in front of the copy, like y0)
        ch.ysm = 0u;
        if constexpr (y_in_smem<R>())
            ch.ysm = (unsigned)__cvta_generic_to_shared(smem_base()) + (unsigned)gibbs_smem_bytes<R, K, WIDE>(accum, a.n_h)
                     + (unsigned)(((int)a.ybase[slot] - ch.off) * (int)sizeof(R));
        ch.c = reinterpret_cast<const R*>(a.cshift)[slot];
        ch.tab_s = (unsigned)__cvta_generic_to_shared(smem_base()) + (unsigned)(threadIdx.x * sizeof(Entry));
        ch.ring_off = (unsigned)(sizeof(Entry) * K * kGibbsThreads + sizeof(R) * (threadIdx.x >> 5) * (kRing * 4 * K * 32));
        if constexpr (SIG) {
            ch.sld = a.sld;
            ch.s0 = reinterpret_cast<const R*>(a.sigw) + a.sbase[slot] - (long long)ch.off * ch.sld;
            ch.pi_back = a.pi_back;
            ch.opi = nullptr; ch.ocs = 0;
        }
        const int T = ch.T;
        // distinct series per chain: y is streamed from HBM (prefetching passes).  One series: the hot passes read it from the block's
        // shared-memory copy (y_sm_elems > 0); a single series too long for that also takes the streaming passes (stride 1)
        // (fp64 only: measured on C2 x 256, same box — fp64 7.19 -> 7.60e10 with the copy, fp32 2.19 -> 2.13e11: there the extra LDS in
        //  the dependent chain costs more than the 6 instructions per 4 steps it saves, so fp32 keeps reading y through L1)
        constexpr bool kYsm = y_in_smem<R>();
        const bool stream_y = a.yld != 1 || (kYsm && a.y_sm_elems <= 0);
        // padding lanes (T = 0) only exist in the last warp, which therefore counts as ragged
        ch.rag_rows = (int)__reduce_max_sync(0xffffffffu, (unsigned)ch.off);
        const R totS = reinterpret_cast<const R*>(a.totS)[slot], totQ = reinterpret_cast<const R*>(a.totQ)[slot];
        const RngKey key{a.k0, a.k1, a.chain_id[slot]};
        R* __restrict__ const out = reinterpret_cast<R*>(a.out);

        int cnt[K], trans[K][K];
        R Sd[K], Qd[K], sig2[K], mu[K], rho[K];
        Hyper<R, K> hp;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            cnt[i] = a.cnt[i * ns + slot];
            Sd[i] = reinterpret_cast<const R*>(a.Sd)[i * ns + slot];
            Qd[i] = reinterpret_cast<const R*>(a.Qd)[i * ns + slot];
            hp.xi[i] = reinterpret_cast<const R*>(a.xi)[i * ns + slot];
            hp.alpha[i] = (R)a.alpha[i];
            hp.nu[i] = (R)a.nu[i];
            sig2[i] = R(1);
#pragma unroll
            for (int j = 0; j < K; ++j) trans[i][j] = a.trans[(i * K + j) * ns + slot];
        }
        int events = 0;
        SigStats<R, K> sg;                                            // statistics of the signals (SIG)
        R totSm = R(0), totQm = R(0);
        int totM = 0;
        if constexpr (SIG) {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                sg.m[i] = a.cntM[i * ns + slot];
                sg.Sm[i] = reinterpret_cast<const R*>(a.Sm)[i * ns + slot];
                sg.Qm[i] = reinterpret_cast<const R*>(a.Qm)[i * ns + slot];
            }
            sg.k1 = (R)(1.0 / (1.0 + a.kappa));
            totSm = reinterpret_cast<const R*>(a.totSm)[slot]; totQm = reinterpret_cast<const R*>(a.totQm)[slot];
            totM = a.totM[slot];
        }

        for (int sw = 0; sw < a.n_sweeps; ++sw) {
            const long long gs = a.sweep0 + sw;
            const uint32_t sweep = (uint32_t)gs;
            // ---- 1. conjugate draws (update_μσ! :231-336 with β of this sweep — quirk Q2 —, update_ρ!, update_A!)
#pragma unroll
            for (int i = 0; i < K; ++i) hp.beta[i] = (R)(gs == 0 ? a.beta0[i] : a.beta[i]);
            events += draw_params<R, K, SIG>(cnt, Sd, Qd, trans, ch.c, hp, key, sweep, sig2, mu, rho, ch.A, &sg);

            // ---- relabelling ranks (:501) and whether this sweep is a saved draw: both passes need them
            ranks_of<R, K>(mu, ch.rank);
            const long long draw_idx = gs - a.burnin;
            const bool save = (draw_idx >= 0) && (T > 0);
            if (accum && save && ch.n_hi > 0) {
                // A^h mu for every requested horizon (chain labels), read back by the in-sample forecasts of the backward (SMOOTH) or forward (filtered means) pass
                R* mh = reinterpret_cast<R*>(smem_base() + ch.mh_off);
                R v[K];
#pragma unroll
                for (int s = 0; s < K; ++s) v[s] = mu[s];
                int h = 0;
                for (int j = 0; j < a.n_h; ++j) {
                    for (; h < a.h_sorted[j]; ++h) {
                        R nv[K];
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            R acc = ch.A[r][0] * v[0];
#pragma unroll
                            for (int s = 1; s < K; ++s) acc = fma(ch.A[r][s], v[s], acc);
                            nv[r] = acc;
                        }
#pragma unroll
                        for (int s = 0; s < K; ++s) v[s] = nv[s];
                    }
#pragma unroll
                    for (int s = 0; s < K; ++s) mh[(a.h_slot[j] * K + s) * kGibbsThreads] = v[s];
                }
            }

            // ---- 2. forward filter
            Emission<R, K> em;
            em.prepare(mu, sig2);
            R pf[K], ll;
            {
                Vec rv;
#pragma unroll
                for (int s = 0; s < K; ++s) rv.v[s] = rho[s];
                FwdOut fo;
                bool facc_done = false;
                if constexpr (!SMOOTH && !SIG) {
                    if (filt && save) {                              // (warp-uniform up to padding lanes, whose T = 0 keeps them out of every row)
                        // (the per-step checked form: a row that would be NaN must never reach the sums, so there is no re-run)
                        fo = stream_y ? forward_pass<true, true, true, true>(ch, em, rv) : forward_pass<true, true, false, true>(ch, em, rv);
                        facc_done = true;
                    }
                }
                if (!facc_done)
                    fo = stream_y ? (ch.rag_rows > 0 ? forward_pass<true, false, true>(ch, em, rv) : forward_pass<false, false, true>(ch, em, rv))
                                  : (ch.rag_rows > 0 ? forward_pass<true, false, false, false, kYsm>(ch, em, rv) : forward_pass<false, false, false, false, kYsm>(ch, em, rv));
                R chk = fo.pf.v[0];
#pragma unroll
                for (int s = 1; s < K; ++s) chk += fo.pf.v[s];
                // a zero / non-finite normaliser anywhere leaves a NaN in the last row: redo the pass with per-step handling
                if (__builtin_expect(!facc_done && T > 0 && !(chk > R(0.5) && chk < R(2)), 0)) fo = forward_pass<true, true>(ch, em, rv);
#pragma unroll
                for (int s = 0; s < K; ++s) pf[s] = fo.pf.v[s];
                ll = fo.ll;
                events += fo.events;
            }
            // pf now holds pif[T,:] in chain labels

            // ---- 3. relabel (:501-513) and emit the draw in increasing-μ order
            if (save) {
                const size_t i = (size_t)(draw_idx - a.draw0);
                const size_t cs = (size_t)a.chunk * ns;             // stride between fields
                R* o = out + i * ns + slot;
                if constexpr (SIG) { ch.opi = o + (size_t)(2 * K + K * K) * cs; ch.ocs = cs; }
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    o[(size_t)(ch.rank[s]) * cs] = mu[s];
                    o[(size_t)(K + ch.rank[s]) * cs] = sig2[s];
                    o[(size_t)(2 * K + K * K + ch.rank[s]) * cs] = pf[s];  // pib[N,:] = pif[N,:]  (:448)
#pragma unroll
                    for (int r = 0; r < K; ++r) o[(size_t)(2 * K + ch.rank[s] * K + ch.rank[r]) * cs] = ch.A[r][s];
                }
                // forecasts pib_T' A^h μ for every requested horizon in one pass over h (:658-667, :858-862)
                const int f0 = 3 * K + K * K;
                R v[K];
#pragma unroll
                for (int s = 0; s < K; ++s) v[s] = pf[s];
                int h = 0;
                for (int j = 0; j < a.n_h; ++j) {
                    for (; h < a.h_sorted[j]; ++h) {
                        R nv[K];
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            R acc = v[0] * ch.A[0][s];
#pragma unroll
                            for (int r = 1; r < K; ++r) acc = fma(v[r], ch.A[r][s], acc);
                            nv[s] = acc;
                        }
#pragma unroll
                        for (int s = 0; s < K; ++s) v[s] = nv[s];
                    }
                    R f = v[0] * mu[0];
#pragma unroll
                    for (int s = 1; s < K; ++s) f = fma(v[s], mu[s], f);
                    const R yr = reinterpret_cast<const R*>(smem_base() + yf_off)[j * kGibbsThreads];
                    o[(size_t)(f0 + 2 * a.h_slot[j]) * cs] = f;
                    o[(size_t)(f0 + 2 * a.h_slot[j] + 1) * cs] = f - yr;
                }
                if (LOGLIK) o[(size_t)(f0 + 2 * a.n_h) * cs] = ll;
            }


            // ---- 4. backward pass
            Back b;
#pragma unroll
            for (int x = 0; x < K; ++x) {                           // this thread's private selection table
                Entry en;
#pragma unroll
                for (int r = 0; r < K; ++r) en.a[r] = ch.A[r][x];
                en.inc = (Row)1 << (Pack::kBits * x);
                smem_tab[x * kGibbsThreads + threadIdx.x] = en;
            }
            int xN;
            typename std::conditional<Pack::kFlush, TransAcc, NoAcc>::type flushed;
            {
                Vec pv;
#pragma unroll
                for (int s = 0; s < K; ++s) pv.v[s] = pf[s];
                BackOut bo;
                if (SMOOTH) {                                        // accumulates into memory: run gated in place
                    bo = backward_pass<true, true>(ch, pv, key, sweep, a.flags, save);
                } else {
                    bool done_b = false;
                    if constexpr (SIG) {
                        if (a.all_signal) {                          // every time step a signal: no observation statistics (warp-uniform)
                            bo = stream_y ? (ch.rag_rows > 0 ? backward_pass<true, false, true, true>(ch, pv, key, sweep, a.flags, save)
                                                             : backward_pass<false, false, true, true>(ch, pv, key, sweep, a.flags, save))
                                          : (ch.rag_rows > 0 ? backward_pass<true, false, false, true, kYsm>(ch, pv, key, sweep, a.flags, save)
                                                             : backward_pass<false, false, false, true, kYsm>(ch, pv, key, sweep, a.flags, save));
                            done_b = true;
                        }
                    }
                    if (!done_b)
                    bo = stream_y ? (ch.rag_rows > 0 ? backward_pass<true, false, true>(ch, pv, key, sweep, a.flags, save)
                                               : backward_pass<false, false, true>(ch, pv, key, sweep, a.flags, save))
                                  : (ch.rag_rows > 0 ? backward_pass<true, false, false, false, kYsm>(ch, pv, key, sweep, a.flags, save)
                                               : backward_pass<false, false, false, false, kYsm>(ch, pv, key, sweep, a.flags, save));
                    // quirk Q5 fired somewhere: redo the pass exactly (counter-based RNG: identical draws otherwise)
                    // (warp-uniform: the pass stages rows through the warp's cp.async ring and synchronises the warp, so every
                    //  lane must take part; lanes that had no Q5 case get identical results from the gated pass)
                    if (__builtin_expect(__any_sync(0xffffffffu, bo.bad), 0)) bo = backward_pass<true, true>(ch, pv, key, sweep, a.flags, save);
                }
                b = bo.b;
                xN = bo.xN;
                if (bo.lost) events += 1 << 20;                      // a bulk copy never completed: the sweep used stale rows — reported, never silent
                if constexpr (Pack::kFlush) {
#pragma unroll
                    for (int q = 0; q < K * K; ++q) flushed.n[q] = bo.n[q];
                }
            }

            // ---- unpack the statistics for the next sweep's draws
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    trans[i][j] = b.tr.get(i, j);
                    if constexpr (Pack::kFlush) trans[i][j] += flushed.n[i * K + j];
                }
            {
                // occupation counts from the transition counts: n_i = sum_j n_ij + [X_N = i]
                R sS = R(0), sQ = R(0);
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    int n = (T > 0 && xN == i) ? 1 : 0;
#pragma unroll
                    for (int j = 0; j < K; ++j) n += trans[i][j];
                    cnt[i] = n;
                }
#pragma unroll
                for (int i = 0; i < K - 1; ++i) { Sd[i] = b.Sd[i]; Qd[i] = b.Qd[i]; sS += b.Sd[i]; sQ += b.Qd[i]; }
                Sd[K - 1] = totS - sS; Qd[K - 1] = totQ - sQ;
                if constexpr (SIG) {
                    // cnt so far counts every time step of a state: split into observations and signals
                    R mS = R(0), mQ = R(0);
                    int mM = 0;
#pragma unroll
                    for (int i = 0; i < K - 1; ++i) {
                        sg.m[i] = b.Mi[i]; sg.Sm[i] = b.Sm[i]; sg.Qm[i] = b.Qm[i];
                        mM += b.Mi[i]; mS += b.Sm[i]; mQ += b.Qm[i];
                    }
                    sg.m[K - 1] = totM - mM; sg.Sm[K - 1] = totSm - mS; sg.Qm[K - 1] = totQm - mQ;
                    if (sg.m[K - 1] == 0) { sg.Sm[K - 1] = R(0); sg.Qm[K - 1] = R(0); }
#pragma unroll
                    for (int i = 0; i < K; ++i) cnt[i] -= sg.m[i];
                }
                if (cnt[K - 1] == 0) { Sd[K - 1] = R(0); Qd[K - 1] = R(0); }
            }
        }

        // ---- store the chain state for the next launch
#pragma unroll
        for (int i = 0; i < K; ++i) {
            a.cnt[i * ns + slot] = cnt[i];
            reinterpret_cast<R*>(a.Sd)[i * ns + slot] = Sd[i];
            reinterpret_cast<R*>(a.Qd)[i * ns + slot] = Qd[i];
#pragma unroll
            for (int j = 0; j < K; ++j) a.trans[(i * K + j) * ns + slot] = trans[i][j];
            if constexpr (SIG) {
                a.cntM[i * ns + slot] = sg.m[i];
                reinterpret_cast<R*>(a.Sm)[i * ns + slot] = sg.Sm[i];
                reinterpret_cast<R*>(a.Qm)[i * ns + slot] = sg.Qm[i];
            }
        }
        a.events[slot] += events;
    }
};

template <typename R, int K, bool WIDE> __host__ __device__ constexpr size_t gibbs_smem_bytes(bool smooth, int n_h) {
    return sizeof(GibbsEntry<R, K, WIDE>) * K * kGibbsThreads                                    // selection tables
           + (HMC_ASYNC ? sizeof(R) * (size_t)(kGibbsThreads / 32) * gibbs_ring_stages<R, K>() * 4 * K * 32 : 0)      // cp.async rings
           + (HMC_TMA ? (size_t)(kGibbsThreads / 32) * gibbs_ring_stages<R, K>() * 8 : 0)                              // their mbarriers (HMC_TMA)
           + (smooth ? sizeof(R) * (size_t)n_h * K * kGibbsThreads : 0)                           // A^h mu (in-sample forecasts)
           + sizeof(R) * (size_t)n_h * kGibbsThreads;                                             // realised y at end+h, per thread
}

template <typename R, int K, bool SMOOTH, bool LOGLIK, bool WIDE, bool SIG = false>
__global__ void __launch_bounds__(kGibbsThreads, gibbs_min_blocks<R, K, SIG>()) gibbs_sweeps_kernel(const GibbsArgs a) {
    using W = GibbsWarp<R, K, SMOOTH, LOGLIK, WIDE, SIG>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename W::Entry* tab = reinterpret_cast<typename W::Entry*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (kGibbsThreads / 32) + (threadIdx.x >> 5);
    if (y_in_smem<R>() && a.y_sm_elems > 0) {                        // ONE short series (fp64): every block keeps a copy (see YRef)
        R* ysm = reinterpret_cast<R*>(smem_raw + gibbs_smem_bytes<R, K, WIDE>((a.flags & (8u | 64u)) != 0u, a.n_h));
        const R* ysrc = reinterpret_cast<const R*>(a.y);
        for (int i = threadIdx.x; i < a.y_sm_elems; i += kGibbsThreads) ysm[i] = __ldg(ysrc + i);
        __syncthreads();
    }
    if (g >= a.n_tasks) return;
    W::run(a, a.task0 + g * a.task_stride, lane, tab);
}

// host-side launcher, instantiated once per (R, K) in gibbs_inst.cu (one translation unit each, built in parallel)
struct GibbsLaunch { unsigned flags; int max_T; int sm_count; int n_h; bool sig; };
template <typename R, int K> cudaError_t launch_gibbs(const GibbsLaunch& cfg, const GibbsArgs& a, cudaStream_t st);
// how many persistent warps of this kernel variant fit on the device at once
template <typename R, int K> int gibbs_capacity_warps(const GibbsLaunch& cfg);

}  // namespace hmc
