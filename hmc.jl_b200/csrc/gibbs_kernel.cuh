// gibbs_kernel.cuh — the fused Gibbs sweep kernel: one thread owns one (window x chain) Markov chain and runs
// whole sweeps on-device (src/Hmc.jl:486-515 gibbssweep!, looped as in gibbssample! :545-560):
//   conjugate draws -> forward filter (pif spilled to HBM, coalesced [t][k][lane]) -> relabel/emit the draw ->
//   backward state sampling fused with the next sweep's sufficient statistics (X never touches memory) ->
//   optional backward smoother accumulation.
// A warp = 32 chains in lockstep; lanes of a warp share one contiguous pif tile so every global access is a
// full 128-byte (fp32) / 256-byte (fp64) line.
#pragma once
#include "hmm_device.cuh"

namespace hmc {

constexpr int kMaxH = 16;
constexpr int kGibbsThreads = 128;

struct GibbsArgs {
    int n_slots;                 // chains incl. padding, multiple of 32
    const int* T;                // [n_slots] window length (0 = padding lane)
    const long long* ybase;      // [n_slots] element offset of the window's first observation in y
    int yld;                     // stride between consecutive time steps in y (= n_series, time-major)
    const void* y;               // R*
    const int* warp_T;           // [n_slots/32] max T within the warp
    const long long* warp_pi_off;// [n_slots/32] element offset of the warp's pif tile
    void* pi;                    // R*  pif spill: tile[(t*K + k)*32 + lane]
    void* pib_acc;               // R*  same layout, smoothed-probability sums of this chunk (SMOOTH only)
    // chain state between launches
    int* cnt;                    // [K][n_slots]
    int* trans;                  // [K*K][n_slots]
    void* Sd;                    // R* [K][n_slots]  sum (y-c)
    void* Qd;                    // R* [K][n_slots]  sum (y-c)^2
    int* events;                 // [n_slots]
    const void* cshift;          // R* [n_slots]
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

template <typename R, int K, bool SMOOTH, bool LOGLIK>
__global__ void __launch_bounds__(kGibbsThreads) gibbs_sweeps_kernel(const GibbsArgs a) {
    const int slot = blockIdx.x * kGibbsThreads + threadIdx.x;
    if (slot >= a.n_slots) return;
    const int lane = threadIdx.x & 31, warp = slot >> 5;
    const int ns = a.n_slots;
    const int T = a.T[slot];
    const int Tw = a.warp_T[warp];
    R* __restrict__ pi = reinterpret_cast<R*>(a.pi) + a.warp_pi_off[warp] + lane;
    R* __restrict__ pacc = SMOOTH ? reinterpret_cast<R*>(a.pib_acc) + a.warp_pi_off[warp] + lane : nullptr;
    const R* __restrict__ y = reinterpret_cast<const R*>(a.y) + a.ybase[slot];
    const size_t yld = (size_t)a.yld;
    const R c = reinterpret_cast<const R*>(a.cshift)[slot];
    const RngKey key{a.k0, a.k1, a.chain_id[slot]};
    R* __restrict__ out = reinterpret_cast<R*>(a.out);

    int cnt[K], trans[K][K];
    R Sd[K], Qd[K], sig2[K], mu[K], rho[K], A[K][K];
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
        events += draw_params<R, K>(cnt, Sd, Qd, trans, c, hp, key, sweep, sig2, mu, rho, A);

        // ---- 2. forward filter (forwardupdate_P! :371-440), pif_t stored for the backward pass
        Emission<R, K> em;
        em.prepare(mu, sig2);
        R pf[K];
#pragma unroll
        for (int s = 0; s < K; ++s) pf[s] = rho[s];                  // t = 1 uses ρ (:390)
        R ll = R(0);
        for (int t = 0; t < Tw; ++t) {
            if (t < T) {
                const R yt = y[(size_t)t * yld];
                R e[K];
                const R m2 = em.eval(yt, e);
                bool ok;
                const R tot = forward_step<R, K>(A, e, pf, ok);
                if (!ok) ++events;
                if (LOGLIK) {
                    if (sizeof(R) == 4) ll += (Real<float>::lg2((float)tot) + (float)m2) * 0.6931471805599453f;
                    else ll += (R)log((double)tot);
                }
#pragma unroll
                for (int s = 0; s < K; ++s) pi[(size_t)(t * K + s) * 32] = pf[s];
            }
        }
        // pf now holds pif[T,:] in chain labels

        // ---- 3. relabel (:501-513) and emit the draw in increasing-μ order
        int rank[K];
        ranks_of<R, K>(mu, rank);
        const long long draw = gs - a.burnin;
        const bool save = (draw >= 0) && (T > 0);
        if (save) {
            const size_t i = (size_t)(draw - a.draw0);
            const size_t cs = (size_t)a.chunk * ns;                 // stride between fields
            R* o = out + i * ns + slot;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                o[(size_t)(rank[s]) * cs] = mu[s];
                o[(size_t)(K + rank[s]) * cs] = sig2[s];
                o[(size_t)(2 * K + K * K + rank[s]) * cs] = pf[s];  // pib[N,:] = pif[N,:]  (:448)
#pragma unroll
                for (int r = 0; r < K; ++r) o[(size_t)(2 * K + rank[s] * K + rank[r]) * cs] = A[r][s];
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
                        R acc = v[0] * A[0][s];
#pragma unroll
                        for (int r = 1; r < K; ++r) acc = fma(v[r], A[r][s], acc);
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

        // ---- 4. backward state sampling (update_X! :459-484) fused with the next sweep's statistics
        //         (update_μσ! :254-258/:291-294 and update_A! :362-365) and, optionally, backwardupdate_P! (:442-457)
#pragma unroll
        for (int i = 0; i < K; ++i) {
            cnt[i] = 0; Sd[i] = R(0); Qd[i] = R(0);
#pragma unroll
            for (int j = 0; j < K; ++j) trans[i][j] = 0;
        }
        int xn = 0;
        R gate = R(1);
        R Acol[K], pb[K];
#pragma unroll
        for (int s = 0; s < K; ++s) { Acol[s] = R(0); pb[s] = R(0); }
        uint4 w = make_uint4(0, 0, 0, 0);
        for (int t = Tw - 1; t >= 0; --t) {
            if ((t & 3) == 3 || t == Tw - 1) w = rng_block(key, sweep, (KIND_STATES << 16), (uint32_t)(t >> 2));
            if (t < T) {
                const uint32_t wi = (t & 3) == 0 ? w.x : (t & 3) == 1 ? w.y : (t & 3) == 2 ? w.z : w.w;
                const R u = u01<R>(wi);
                const R yt = y[(size_t)t * yld];
                R pt[K];
                int x;
                if (t == T - 1) {
                    // X[N] ~ Categorical(pif[N,:]) — with quirk Q1 the relabelled row is used with chain labels (:512-514)
                    R pN[K];
#pragma unroll
                    for (int s = 0; s < K; ++s) pt[s] = pf[s];
                    if (a.flags & 1u) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            R vsel = R(0);
#pragma unroll
                            for (int s = 0; s < K; ++s) vsel = (rank[s] == k) ? pf[s] : vsel;
                            pN[k] = vsel;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < K; ++s) pN[s] = pf[s];
                    }
                    x = categorical_unnorm<R, K>(pN, u);
                    if (SMOOTH) {
#pragma unroll
                        for (int s = 0; s < K; ++s) pb[s] = pf[s];   // pib[N,:] = pif[N,:]
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < K; ++s) pt[s] = pi[(size_t)(t * K + s) * 32];
                    x = backward_sample_step<R, K>(Acol, pt, gate, u);
                    if (SMOOTH) smooth_step<R, K>(A, pt, pb);
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int j = 0; j < K; ++j) trans[i][j] += (x == i && xn == j) ? 1 : 0;
                }
                if (SMOOTH && save) {
#pragma unroll
                    for (int s = 0; s < K; ++s) pacc[(size_t)(t * K + rank[s]) * 32] += pb[s];
                }
                const R d = yt - c;
                const R dd = d * d;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const bool m = (x == i);
                    cnt[i] += m ? 1 : 0;
                    Sd[i] += m ? d : R(0);
                    Qd[i] += m ? dd : R(0);
                }
                // prepare the next (earlier) step: column xn of A and gate = pif[t, x]
                xn = x;
                gate = select_k<R, K>(pt, x);
#pragma unroll
                for (int r = 0; r < K; ++r) Acol[r] = select_k<R, K>(A[r], x);
            }
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

}  // namespace hmc
