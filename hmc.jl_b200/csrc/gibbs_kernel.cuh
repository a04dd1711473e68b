// gibbs_kernel.cuh — the fused Gibbs sweep kernel: one thread owns one (window x chain) Markov chain and runs
// whole sweeps on-device (src/Hmc.jl:486-515 gibbssweep!, looped as in gibbssample! :545-560):
//   conjugate draws -> forward filter (pif spilled to HBM, coalesced [row][k][lane]) -> relabel/emit the draw ->
//   backward state sampling fused with the next sweep's sufficient statistics (X never touches memory) ->
//   optional backward smoother accumulation.
//
// Mapping.  A warp = 32 chains in lockstep ("warp task"); lanes of a warp share one contiguous pif tile so every
// global access is a full 128-byte (fp32) / 256-byte (fp64) line.  Warp tasks are sorted by window length
// (longest first) and handed out dynamically through one atomic counter per launch, so the 148 SMs stay balanced
// although T_w varies 6x across the expanding windows.  Windows inside a warp are right-aligned in time: loop row j
// holds t = j - (T_warp - T_lane), which keeps the first backward step and every Philox block warp-uniform.
#pragma once
#include <type_traits>
#include "hmm_device.cuh"

namespace hmc {

constexpr int kMaxH = 16;
constexpr int kGibbsThreads = 128;

struct GibbsArgs {
    int n_slots;                 // chains incl. padding, multiple of 32
    int n_warps;                 // warp tasks = n_slots/32
    int* task_counter;           // zeroed before the launch
    const int* T;                // [n_slots] window length (0 = padding lane)
    const long long* ybase;      // [n_slots] element offset of the window's first observation in y
    int yld;                     // stride between consecutive time steps in y (= n_series, time-major)
    const void* y;               // R*
    const int* warp_T;           // [n_warps] max T within the warp
    const long long* warp_pi_off;// [n_warps] element offset of the warp's pif tile
    void* pi;                    // R*  pif spill: tile[(row*K + k)*32 + lane]
    void* pib_acc;               // R*  same layout, smoothed-probability sums of this chunk (SMOOTH only)
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
    double alpha[8], nu[8], beta0[8], beta[8];
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
};

// Transition counts n_ij of the sampled path, packed: one word per origin state, one bit-field per destination.
template <int K, bool WIDE> struct TransPack {
    using Row = typename std::conditional<WIDE, unsigned long long, unsigned int>::type;
    static constexpr int kBits = (WIDE ? 64 : 32) / K;
    static constexpr long long kMaxT = (1ll << kBits) - 1;
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

template <typename R, int K, bool SMOOTH, bool LOGLIK, bool WIDE>
struct GibbsWarp {
    using Pack = TransPack<K, WIDE>;
    using Row = typename Pack::Row;

    // per-sweep constants of one chain
    struct Chain {
        R A[K][K];
        R c;                      // shift of the sufficient statistics
        int rank[K];              // position of each chain label in increasing-mu order
        int T, Tw, off;           // window length, warp length, right-alignment offset (row j <-> t = j - off)
        long long yld;
        const R* y0;              // row j -> y0[j*yld]
        R* pi0;                   // row j, state k -> pi0[(j*K + k)*32]
        R* pacc0;
    };

    // state carried by the backward pass; the sampled state of the later time step is kept one-hot in (lt[])
    struct Back {
        R Sd[K - 1], Qd[K - 1];   // statistics of states 0..K-2 (the last one follows from the window totals)
        int n[K - 1];
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
    static __device__ __forceinline__ void commit(Back& b, const Chain& ch, const bool (&lt)[K - 1], const R (&pt)[K], R yt, bool first) {
        if (!first) {
#pragma unroll
            for (int i = 0; i < K; ++i) if (is_state(lt, i)) b.tr.row[i] += b.inc;
        }
        const R d = yt - ch.c, dd = d * d;
#pragma unroll
        for (int i = 0; i < K - 1; ++i)
            if (is_state(lt, i)) { b.Sd[i] += d; b.Qd[i] += dd; b.n[i] += 1; }
        Row incs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) incs[i] = (Row)1 << (Pack::kBits * i);
        b.inc = pick<Row>(lt, incs);
        b.gate = pick<R>(lt, pt);
#pragma unroll
        for (int r = 0; r < K; ++r) b.Acol[r] = pick<R>(lt, ch.A[r]);
    }

    // one backward step at loop row j: draw X_t | X_{t+1} (src/Hmc.jl:466-481 in the pif form), update the statistics
    static __device__ __forceinline__ void back_step(Back& b, const Chain& ch, const R* __restrict__ pip, R* __restrict__ pap,
                                                      const R* __restrict__ yp, uint32_t word, bool save) {
        R pt[K];
#pragma unroll
        for (int s = 0; s < K; ++s) pt[s] = pip[s * 32];
        const R yt = *yp;
        R p[K];
#pragma unroll
        for (int r = 0; r < K; ++r) p[r] = pt[r] * b.Acol[r];
        const R u = u01<R>(word);
        bool lt[K - 1];
        draw(p, u, lt);
        if (__builtin_expect(!(b.gate > Real<R>::eps()), 0)) {   // reference: uniform p when total <= eps() (:472-480)
#pragma unroll
            for (int r = 0; r < K; ++r) p[r] = R(1);
            draw(p, u, lt);
        }
        if (SMOOTH) {
            smooth_step<R, K>(ch.A, pt, b.pb);
            if (save) {
#pragma unroll
                for (int s = 0; s < K; ++s) pap[ch.rank[s] * 32] += b.pb[s];
            }
        }
        commit(b, ch, lt, pt, yt, false);
    }

    // forward filter (forwardupdate_P! :371-440); pif rows stored for the backward pass.  CHECKED = per-step handling
    // of a zero / non-finite normaliser (the reference only warns, :435); the hot path runs unchecked and is re-run
    // checked when the final row is not finite (a NaN, once produced, propagates to the last row).
    template <bool RAGGED, bool CHECKED>
    static __device__ __forceinline__ int forward(const Chain& ch, const Emission<R, K>& em, const R (&rho)[K], R (&pf)[K], R& ll) {
        int events = 0;
#pragma unroll
        for (int s = 0; s < K; ++s) pf[s] = rho[s];                  // t = 1 uses ρ (:390)
        ll = R(0);
        const R* yp = ch.y0;
        R* pip = ch.pi0;
        auto step = [&](int j, int u) {
            if (!RAGGED || j >= ch.off) {
                const R yt = yp[u * ch.yld];
                R e[K];
                const R m2 = em.eval(yt, e);
                bool ok;
                const R tot = forward_step<R, K>(ch.A, e, pf, ok);
                if (CHECKED && !ok) {
                    ++events;
#pragma unroll
                    for (int s = 0; s < K; ++s) pf[s] = R(1) / R(K);
                }
                if (LOGLIK) {
                    if (sizeof(R) == 4) ll += (Real<float>::lg2((float)tot) + (float)m2) * 0.6931471805599453f;
                    else ll += (R)log((double)tot);
                }
#pragma unroll
                for (int s = 0; s < K; ++s) pip[(u * K + s) * 32] = pf[s];
            }
        };
        int j = 0;
        for (; j + 3 < ch.Tw; j += 4, yp += 4 * ch.yld, pip += 4 * K * 32) { step(j, 0); step(j + 1, 1); step(j + 2, 2); step(j + 3, 3); }
        for (; j < ch.Tw; ++j, yp += ch.yld, pip += K * 32) step(j, 0);
        return events;
    }
    static __device__ __noinline__ int forward_checked(const Chain& ch, const Emission<R, K>& em, const R (&rho)[K], R (&pf)[K], R& ll) {
        return forward<true, true>(ch, em, rho, pf, ll);
    }

    // backward state sampling (update_X! :459-484) fused with the next sweep's statistics (update_μσ! :254-258/:291-294,
    // update_A! :362-365) and, optionally, backwardupdate_P! (:442-457).
    // The i-th uniform consumed (i = 0 for X[N]) is word i&3 of Philox block i>>2 of this sweep.
    template <bool RAGGED>
    static __device__ __forceinline__ void backward(Back& b, const Chain& ch, const R (&pf)[K], const RngKey& key, uint32_t sweep,
                                                    unsigned flags, bool save) {
        const int Tw = ch.Tw, T = ch.T;
        const R* yp = ch.y0 + (long long)(Tw - 1) * ch.yld;
        const R* pip = ch.pi0 + (size_t)(Tw - 1) * K * 32;
        R* pap = SMOOTH ? ch.pacc0 + (size_t)(Tw - 1) * K * 32 : nullptr;
        uint4 w = rng_block(key, sweep, (KIND_STATES << 16), 0u);
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
            if (SMOOTH) {
#pragma unroll
                for (int s = 0; s < K; ++s) b.pb[s] = pf[s];           // pib[N,:] = pif[N,:]
                if (save) {
#pragma unroll
                    for (int s = 0; s < K; ++s) pap[ch.rank[s] * 32] += pf[s];
                }
            }
            commit(b, ch, lt, pf, *yp, true);
        }
        int i = 1;
        const long long ys = ch.yld;
#define HMC_BACK(word)                                                                   \
    yp -= ys; pip -= K * 32; if (SMOOTH) pap -= K * 32;                                  \
    if (!RAGGED || i < T) back_step(b, ch, pip, pap, yp, (word), save);                  \
    ++i;
        if (Tw > 1) { HMC_BACK(w.y) }
        if (Tw > 2) { HMC_BACK(w.z) }
        if (Tw > 3) { HMC_BACK(w.w) }
        for (; i + 3 < Tw;) {
            w = rng_block(key, sweep, (KIND_STATES << 16), (uint32_t)(i >> 2));
            HMC_BACK(w.x) HMC_BACK(w.y) HMC_BACK(w.z) HMC_BACK(w.w)
        }
        if (i < Tw) {
            w = rng_block(key, sweep, (KIND_STATES << 16), (uint32_t)(i >> 2));
            HMC_BACK(w.x)
            if (i < Tw) { HMC_BACK(w.y) }
            if (i < Tw) { HMC_BACK(w.z) }
        }
#undef HMC_BACK
    }

    static __device__ void run(const GibbsArgs& a, const int warp, const int lane) {
        const int slot = warp * 32 + lane;
        const int ns = a.n_slots;
        Chain ch;
        ch.T = a.T[slot];
        ch.Tw = a.warp_T[warp];
        ch.off = ch.Tw - ch.T;
        ch.yld = a.yld;
        ch.pi0 = reinterpret_cast<R*>(a.pi) + a.warp_pi_off[warp] + lane;
        ch.pacc0 = SMOOTH ? reinterpret_cast<R*>(a.pib_acc) + a.warp_pi_off[warp] + lane : nullptr;
        ch.y0 = reinterpret_cast<const R*>(a.y) + a.ybase[slot] - (long long)ch.off * ch.yld;
        ch.c = reinterpret_cast<const R*>(a.cshift)[slot];
        const int T = ch.T;
        // padding lanes (T = 0) only exist in the last warp, which therefore counts as ragged
        const bool ragged = __any_sync(0xffffffffu, ch.off != 0);
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

        for (int sw = 0; sw < a.n_sweeps; ++sw) {
            const long long gs = a.sweep0 + sw;
            const uint32_t sweep = (uint32_t)gs;
            // ---- 1. conjugate draws (update_μσ! :231-336 with β of this sweep — quirk Q2 —, update_ρ!, update_A!)
#pragma unroll
            for (int i = 0; i < K; ++i) hp.beta[i] = (R)(gs == 0 ? a.beta0[i] : a.beta[i]);
            events += draw_params<R, K>(cnt, Sd, Qd, trans, ch.c, hp, key, sweep, sig2, mu, rho, ch.A);

            // ---- 2. forward filter
            Emission<R, K> em;
            em.prepare(mu, sig2);
            R pf[K], ll;
            if (ragged) forward<true, false>(ch, em, rho, pf, ll);
            else forward<false, false>(ch, em, rho, pf, ll);
            {
                R chk = pf[0];
#pragma unroll
                for (int s = 1; s < K; ++s) chk += pf[s];
                if (__builtin_expect(T > 0 && !(chk > R(0.5) && chk < R(2)), 0)) events += forward_checked(ch, em, rho, pf, ll);
            }
            // pf now holds pif[T,:] in chain labels

            // ---- 3. relabel (:501-513) and emit the draw in increasing-μ order
            ranks_of<R, K>(mu, ch.rank);
            const long long draw_idx = gs - a.burnin;
            const bool save = (draw_idx >= 0) && (T > 0);
            if (save) {
                const size_t i = (size_t)(draw_idx - a.draw0);
                const size_t cs = (size_t)a.chunk * ns;             // stride between fields
                R* o = out + i * ns + slot;
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
                    const R yr = reinterpret_cast<const R*>(a.yfut)[(size_t)a.h_slot[j] * ns + slot];
                    o[(size_t)(f0 + 2 * a.h_slot[j]) * cs] = f;
                    o[(size_t)(f0 + 2 * a.h_slot[j] + 1) * cs] = f - yr;
                }
                if (LOGLIK) o[(size_t)(f0 + 2 * a.n_h) * cs] = ll;
            }

            // ---- 4. backward pass
            Back b;
#pragma unroll
            for (int i = 0; i < K - 1; ++i) { b.Sd[i] = R(0); b.Qd[i] = R(0); b.n[i] = 0; }
            b.tr.clear();
            b.inc = 0; b.gate = R(1);
#pragma unroll
            for (int s = 0; s < K; ++s) { b.Acol[s] = R(0); b.pb[s] = R(0); }
            if (ragged) backward<true>(b, ch, pf, key, sweep, a.flags, save);
            else backward<false>(b, ch, pf, key, sweep, a.flags, save);

            // ---- unpack the statistics for the next sweep's draws
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) trans[i][j] = b.tr.get(i, j);
            {
                R sS = R(0), sQ = R(0);
                int sn = 0;
#pragma unroll
                for (int i = 0; i < K - 1; ++i) { Sd[i] = b.Sd[i]; Qd[i] = b.Qd[i]; cnt[i] = b.n[i]; sS += b.Sd[i]; sQ += b.Qd[i]; sn += b.n[i]; }
                Sd[K - 1] = totS - sS; Qd[K - 1] = totQ - sQ; cnt[K - 1] = T - sn;
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
        }
        a.events[slot] += events;
    }
};

template <typename R, int K, bool SMOOTH, bool LOGLIK, bool WIDE>
__global__ void __launch_bounds__(kGibbsThreads) gibbs_sweeps_kernel(const GibbsArgs a) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(a.task_counter, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_warps) break;
        GibbsWarp<R, K, SMOOTH, LOGLIK, WIDE>::run(a, task, lane);
    }
}

}  // namespace hmc
