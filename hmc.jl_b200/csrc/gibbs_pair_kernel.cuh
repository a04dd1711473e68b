// gibbs_pair_kernel.cuh — fp32 sweep kernel with TWO chains per thread and packed FP32x2 arithmetic.
//
// The thread-per-chain kernel (gibbs_kernel.cuh) is bound by instruction issue, not by HBM or by the FP32 pipes.
// Blackwell (sm_100) has packed two-wide fp32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot, two results), so
// here every thread owns two chains of the same window (same y, same T) and keeps their state in the two halves of
// 64-bit register pairs: the whole forward recursion (emission quadratic, prediction K x K, normalisation) is issued
// once for both chains, the filtered rows are stored/loaded as float2, and loads, loop control, cp.async staging and
// observation reads are shared.  A warp task is therefore 64 chains; everything else (buffers indexed by chain slot,
// Philox streams, statistics, quirks, outputs) is identical to the scalar kernel, whose per-chain pieces are reused.
// Reference lines: see gibbs_kernel.cuh.
#pragma once
#include "gibbs_kernel.cuh"

namespace hmc {

template <int K, bool LOGLIK, bool WIDE>
struct GibbsPair {
    using Pack = TransPack<K, WIDE>;
    using Row = typename Pack::Row;
    struct alignas(16) Entry { float a[K]; Row inc; };

    struct Chain {                 // per-sweep constants of the pair
        f2 A[K][K];
        float c;                   // shift of the sufficient statistics (window mean: shared by the pair)
        int T, Tw, off;
        bool ragged;
        long long yld;
        const float* y0;           // row j -> y0[j*yld]
        float2* pi0;               // row j, state k -> pi0[(j*K + k)*32]   (x: first chain, y: second chain)
        unsigned tab_off[2];       // shared memory (bytes): selection table of each chain
        unsigned ring_off;
        int rank[2][K];
    };
    struct Em { f2 negmu[K], q[K], cc[K]; };
    struct Vec { f2 v[K]; };
    struct Back {                  // per chain, as in the scalar kernel
        float Sd[K - 1], Qd[K - 1];
        Pack tr;
        Row inc;
        float Acol[K];
        float gate;
    };

    static __device__ __forceinline__ bool is_state(const bool (&lt)[K - 1], int i) {
        if (i == 0) return !lt[0];
        if (i == K - 1) return lt[K - 2];
        return lt[i - 1] && !lt[i];
    }
    static __device__ __forceinline__ float pick(const bool (&lt)[K - 1], const float (&v)[K]) {
        float r = v[0];
#pragma unroll
        for (int i = 1; i < K; ++i) r = lt[i - 1] ? v[i] : r;
        return r;
    }
    static __device__ __forceinline__ void draw(const float (&p)[K], float u, bool (&lt)[K - 1]) {
        float cum[K];
        cum[0] = p[0];
#pragma unroll
        for (int i = 1; i < K; ++i) cum[i] = cum[i - 1] + p[i];
        const float thr = u * cum[K - 1];
#pragma unroll
        for (int i = 0; i < K - 1; ++i) lt[i] = cum[i] < thr;
    }
    static __device__ __forceinline__ void commit(Back& b, float c, unsigned tab_off, const bool (&lt)[K - 1], const float (&pt)[K],
                                                  float yt, bool first) {
        if (!first) {
#pragma unroll
            for (int i = 0; i < K; ++i) if (is_state(lt, i)) b.tr.row[i] += b.inc;
        }
        const float d = yt - c, dd = d * d;
#pragma unroll
        for (int i = 0; i < K - 1; ++i)
            if (is_state(lt, i)) { b.Sd[i] += d; b.Qd[i] += dd; }
        const Entry* e = reinterpret_cast<const Entry*>(smem_base() + tab_off);
#pragma unroll
        for (int i = 1; i < K; ++i) if (lt[i - 1]) e += kGibbsThreads;
        const Entry en = *e;
        b.inc = en.inc;
#pragma unroll
        for (int r = 0; r < K; ++r) b.Acol[r] = en.a[r];
        b.gate = pick(lt, pt);
    }

    // ---------------------------------------------------------------- forward pass, both chains per instruction
    struct FwdOut { Vec pf; float ll[2]; int events; };
    template <bool RAGGED, bool CHECKED>
    static __device__ __noinline__ FwdOut forward_pass(const Chain ch, const Em em, const Vec rho) {
        FwdOut o;
        f2 (&pf)[K] = o.pf.v;
#pragma unroll
        for (int s = 0; s < K; ++s) pf[s] = rho.v[s];
        float ll0 = 0.f, ll1 = 0.f;
        int events = 0;
        const long long yld = ch.yld;
        const float* yp = ch.y0;
        float2* pip = ch.pi0;
        auto step = [&](int j, int u) {
            if (!RAGGED || j >= ch.off) {
                const f2 yt = splat2(ld_ro(yp + u * yld));
                f2 l[K];
#pragma unroll
                for (int s = 0; s < K; ++s) { const f2 d = yt + em.negmu[s]; l[s] = fma2(d * d, em.q[s], em.cc[s]); }
                float mx = l[0].v.x, my = l[0].v.y;
#pragma unroll
                for (int s = 1; s < K; ++s) { mx = fmaxf(mx, l[s].v.x); my = fmaxf(my, l[s].v.y); }
                const f2 nm = mk2(-mx, -my);
                f2 q[K];
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const f2 a = l[s] + nm;
                    const f2 e = mk2(Real<float>::ex2(a.v.x), Real<float>::ex2(a.v.y));
                    f2 pred = pf[0] * ch.A[0][s];
#pragma unroll
                    for (int r = 1; r < K; ++r) pred = fma2(pf[r], ch.A[r][s], pred);
                    q[s] = pred * e;
                }
                f2 tot = q[0];
#pragma unroll
                for (int s = 1; s < K; ++s) tot = tot + q[s];
                const f2 inv = mk2(Real<float>::rcp(tot.v.x), Real<float>::rcp(tot.v.y));
#pragma unroll
                for (int s = 0; s < K; ++s) pf[s] = q[s] * inv;
                if (CHECKED) {
                    if (!((tot.v.x > 0.f) && (tot.v.x < 3.0e38f))) {
                        ++events;
#pragma unroll
                        for (int s = 0; s < K; ++s) pf[s].v.x = 1.f / K;
                    }
                    if (!((tot.v.y > 0.f) && (tot.v.y < 3.0e38f))) {
                        ++events;
#pragma unroll
                        for (int s = 0; s < K; ++s) pf[s].v.y = 1.f / K;
                    }
                }
                if (LOGLIK) {
                    ll0 += (Real<float>::lg2(tot.v.x) + mx) * 0.6931471805599453f;
                    ll1 += (Real<float>::lg2(tot.v.y) + my) * 0.6931471805599453f;
                }
#pragma unroll
                for (int s = 0; s < K; ++s) st_stream(pip + (u * K + s) * 32, pf[s].v);
            }
        };
        int j = 0;
        for (; j + 3 < ch.Tw; j += 4, yp += 4 * yld, pip += 4 * K * 32) { step(j, 0); step(j + 1, 1); step(j + 2, 2); step(j + 3, 3); }
        for (; j < ch.Tw; ++j, yp += yld, pip += K * 32) step(j, 0);
        o.ll[0] = ll0; o.ll[1] = ll1;
        o.events = events;
        return o;
    }

    // ---------------------------------------------------------------- backward pass: shared loads/staging, per-chain sampling
    struct BackOut { Back b[2]; int xN[2]; bool bad; };
    template <bool RAGGED, bool GATED>
    static __device__ __noinline__ BackOut backward_pass(const Chain ch, const Vec pf_in, const RngKey key0, const RngKey key1,
                                                         const uint32_t sweep, const unsigned flags) {
        BackOut o;
        bool bad = false;
        const int Tw = ch.Tw, T = ch.T;
        const long long ys = ch.yld;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            Back& b = o.b[c];
#pragma unroll
            for (int i = 0; i < K - 1; ++i) { b.Sd[i] = 0.f; b.Qd[i] = 0.f; }
            b.tr.clear();
            b.inc = 0; b.gate = 1.f;
#pragma unroll
            for (int s = 0; s < K; ++s) b.Acol[s] = 0.f;
            o.xN[c] = 0;
        }
        const float* yp = ch.y0 + (long long)(Tw - 1) * ys;
        const float2* pip = ch.pi0 + (size_t)(Tw - 1) * K * 32;
        uint4 w0 = rng_block_states(key0, sweep, 0u);
        uint4 w1 = rng_block_states(key1, sweep, 0u);
        // one chain's step: draw X_t | X_{t+1}; quirk Q5 only recorded unless GATED (see gibbs_kernel.cuh)
        auto chain_step = [&](Back& b, unsigned tab_off, const float (&pt)[K], float yt, uint32_t word) {
            float p[K];
#pragma unroll
            for (int r = 0; r < K; ++r) p[r] = pt[r] * b.Acol[r];
            if (GATED) {
                if (!(b.gate > Real<float>::eps())) {
#pragma unroll
                    for (int r = 0; r < K; ++r) p[r] = 1.f;
                }
            } else {
                bad = bad || !(b.gate > Real<float>::eps());
            }
            bool lt[K - 1];
            draw(p, u01<float>(word), lt);
            commit(b, ch.c, tab_off, lt, pt, yt, false);
        };
        auto pair_step = [&](const float2 (&pt2)[K], float yt, uint32_t wa, uint32_t wb) {
            float pa[K], pb[K];
#pragma unroll
            for (int s = 0; s < K; ++s) { pa[s] = pt2[s].x; pb[s] = pt2[s].y; }
            chain_step(o.b[0], ch.tab_off[0], pa, yt, wa);
            chain_step(o.b[1], ch.tab_off[1], pb, yt, wb);
        };
        if (T > 0) {
            // X[N] ~ Categorical(pif[N,:]); with quirk Q1 the relabelled row is used with chain labels (:512-514)
            const float yN = ld_ro(yp);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float pfc[K], pN[K];
#pragma unroll
                for (int s = 0; s < K; ++s) pfc[s] = c == 0 ? pf_in.v[s].v.x : pf_in.v[s].v.y;
                if (flags & 1u) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        float vsel = 0.f;
#pragma unroll
                        for (int s = 0; s < K; ++s) vsel = (ch.rank[c][s] == k) ? pfc[s] : vsel;
                        pN[k] = vsel;
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < K; ++s) pN[s] = pfc[s];
                }
                bool lt[K - 1];
                draw(pN, u01<float>(c == 0 ? w0.x : w1.x), lt);
                int xN = 0;
#pragma unroll
                for (int i = 0; i < K - 1; ++i) xN += lt[i] ? 1 : 0;
                o.xN[c] = xN;
                commit(o.b[c], ch.c, ch.tab_off[c], lt, pfc, yN, true);
            }
        }
        int i = 1;
        auto load_row = [&](int u, float2 (&pt)[K], float& yt) {
#pragma unroll
            for (int s = 0; s < K; ++s) pt[s] = ld_stream(pip + (s - (u + 1) * K) * 32);
            yt = (!RAGGED || i + u < T) ? ld_ro(yp - (u + 1) * ys) : 0.f;
        };
#define HMC_PBACK(u, WA, WB, PT, YT) if (!RAGGED || i + (u) < T) pair_step(PT, YT, WA, WB);
        {
            float2 p0[K], p1[K], p2[K];
            float y0 = 0.f, y1 = 0.f, y2 = 0.f;
            if (Tw > 1) load_row(0, p0, y0);
            if (Tw > 2) load_row(1, p1, y1);
            if (Tw > 3) load_row(2, p2, y2);
            if (Tw > 1) { HMC_PBACK(0, w0.y, w1.y, p0, y0) }
            if (Tw > 2) { HMC_PBACK(1, w0.z, w1.z, p1, y1) }
            if (Tw > 3) { HMC_PBACK(2, w0.w, w1.w, p2, y2) }
            const int done = Tw > 3 ? 3 : Tw - 1;
            i += done; yp -= done * ys; pip -= (size_t)done * K * 32;
        }
        {
            // groups of 4 rows (4*K*32 float2, contiguous) staged global -> shared with cp.async, kRing-1 groups ahead
            constexpr int kGroupElems = 4 * K * 32;
            constexpr int kChunksPerLane = (int)(kGroupElems * sizeof(float2) / 16 / 32);
            const int lane = threadIdx.x & 31;
            float2* const ring = reinterpret_cast<float2*>(smem_base() + ch.ring_off);
            const int n_groups = (Tw - i) / 4;
            const float2* gsrc = ch.pi0 - lane + (long long)(Tw - 8) * K * 32;
            auto issue = [&](int g) {
                if (g < n_groups) {
                    const char* src = reinterpret_cast<const char*>(gsrc - (long long)g * kGroupElems);
                    char* dst = reinterpret_cast<char*>(ring + (g % kRing) * kGroupElems);
#pragma unroll
                    for (int m = 0; m < kChunksPerLane; ++m) cp_async16(dst + (lane + 32 * m) * 16, src + (lane + 32 * m) * 16);
                }
                cp_async_commit();
            };
#pragma unroll
            for (int g = 0; g < kRing - 1; ++g) issue(g);
            for (int g = 0; g < n_groups; ++g, i += 4, yp -= 4 * ys, pip -= 4 * K * 32) {
                issue(g + kRing - 1);
                cp_async_wait<kRing - 1>();
                __syncwarp();
                const float2* st = ring + (g % kRing) * kGroupElems + lane;
                float2 c0[K], c1[K], c2[K], c3[K];
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    c0[s] = st[(3 * K + s) * 32]; c1[s] = st[(2 * K + s) * 32]; c2[s] = st[(1 * K + s) * 32]; c3[s] = st[s * 32];
                }
                const float y0 = (!RAGGED || i + 0 < T) ? ld_ro(yp - 1 * ys) : 0.f;
                const float y1 = (!RAGGED || i + 1 < T) ? ld_ro(yp - 2 * ys) : 0.f;
                const float y2 = (!RAGGED || i + 2 < T) ? ld_ro(yp - 3 * ys) : 0.f;
                const float y3 = (!RAGGED || i + 3 < T) ? ld_ro(yp - 4 * ys) : 0.f;
                w0 = rng_block_states(key0, sweep, (uint32_t)(i >> 2));
                w1 = rng_block_states(key1, sweep, (uint32_t)(i >> 2));
                HMC_PBACK(0, w0.x, w1.x, c0, y0) HMC_PBACK(1, w0.y, w1.y, c1, y1)
                HMC_PBACK(2, w0.z, w1.z, c2, y2) HMC_PBACK(3, w0.w, w1.w, c3, y3)
                __syncwarp();
            }
            cp_async_wait<0>();
        }
        if (i < Tw) {
            w0 = rng_block_states(key0, sweep, (uint32_t)(i >> 2));
            w1 = rng_block_states(key1, sweep, (uint32_t)(i >> 2));
            float2 p0[K], p1[K], p2[K];
            float y0 = 0.f, y1 = 0.f, y2 = 0.f;
            load_row(0, p0, y0);
            if (i + 1 < Tw) load_row(1, p1, y1);
            if (i + 2 < Tw) load_row(2, p2, y2);
            HMC_PBACK(0, w0.x, w1.x, p0, y0)
            if (i + 1 < Tw) { HMC_PBACK(1, w0.y, w1.y, p1, y1) }
            if (i + 2 < Tw) { HMC_PBACK(2, w0.z, w1.z, p2, y2) }
        }
#undef HMC_PBACK
        o.bad = bad;
        return o;
    }

    // ---------------------------------------------------------------- one warp task = 64 chains (32 pairs)
    static __device__ void run(const GibbsArgs& a, const int task, const int lane, Entry* smem_tab) {
        const int ns = a.n_slots;
        const int slot0 = task * 64 + 2 * lane;                     // the pair: slots slot0, slot0 + 1 (same window)
        Chain ch;
        ch.T = a.T[slot0];
        ch.Tw = a.warp_T[task];
        ch.off = ch.Tw - ch.T;
        ch.yld = a.yld;
        ch.pi0 = reinterpret_cast<float2*>(reinterpret_cast<float*>(a.pi) + a.warp_pi_off[task]) + lane;
        ch.y0 = reinterpret_cast<const float*>(a.y) + a.ybase[slot0] - (long long)ch.off * ch.yld;
        ch.c = reinterpret_cast<const float*>(a.cshift)[slot0];
        ch.tab_off[0] = (unsigned)(threadIdx.x * sizeof(Entry));
        ch.tab_off[1] = (unsigned)((K * kGibbsThreads + threadIdx.x) * sizeof(Entry));
        ch.ring_off = (unsigned)(sizeof(Entry) * 2 * K * kGibbsThreads + sizeof(float2) * (threadIdx.x >> 5) * (kRing * 4 * K * 32));
        ch.ragged = __any_sync(0xffffffffu, ch.off != 0);
        const int T = ch.T;
        const float totS = reinterpret_cast<const float*>(a.totS)[slot0], totQ = reinterpret_cast<const float*>(a.totQ)[slot0];
        const RngKey key[2] = {{a.k0, a.k1, a.chain_id[slot0]}, {a.k0, a.k1, a.chain_id[slot0 + 1]}};
        float* __restrict__ const out = reinterpret_cast<float*>(a.out);

        int cnt[2][K], trans[2][K][K];
        float Sd[2][K], Qd[2][K], sig2[2][K], mu[2][K], rho[2][K], A[2][K][K];
        Hyper<float, K> hp;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int slot = slot0 + c;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                cnt[c][i] = a.cnt[i * ns + slot];
                Sd[c][i] = reinterpret_cast<const float*>(a.Sd)[i * ns + slot];
                Qd[c][i] = reinterpret_cast<const float*>(a.Qd)[i * ns + slot];
                sig2[c][i] = 1.f;
#pragma unroll
                for (int j = 0; j < K; ++j) trans[c][i][j] = a.trans[(i * K + j) * ns + slot];
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            hp.xi[i] = reinterpret_cast<const float*>(a.xi)[i * ns + slot0];      // same window -> same prior mean
            hp.alpha[i] = (float)a.alpha[i];
            hp.nu[i] = (float)a.nu[i];
        }
        int events[2] = {0, 0};

        for (int sw = 0; sw < a.n_sweeps; ++sw) {
            const long long gs = a.sweep0 + sw;
            const uint32_t sweep = (uint32_t)gs;
            // ---- 1. conjugate draws, per chain
#pragma unroll
            for (int i = 0; i < K; ++i) hp.beta[i] = (float)(gs == 0 ? a.beta0[i] : a.beta[i]);
#pragma unroll
            for (int c = 0; c < 2; ++c)
                events[c] += draw_params<float, K>(cnt[c], Sd[c], Qd[c], trans[c], ch.c, hp, key[c], sweep, sig2[c], mu[c], rho[c], A[c]);

            // ---- 2. forward filter, packed
            Em em;
            Vec rv;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                Emission<float, K> e0, e1;     // only used for the formulas of q and c
                (void)e0; (void)e1;
                em.negmu[s] = mk2(-mu[0][s], -mu[1][s]);
                em.q[s] = mk2(-0.72134752044448170368f / sig2[0][s], -0.72134752044448170368f / sig2[1][s]);
                em.cc[s] = mk2(-0.5f * Real<float>::lg2(6.283185307179586f * sig2[0][s]), -0.5f * Real<float>::lg2(6.283185307179586f * sig2[1][s]));
                rv.v[s] = mk2(rho[0][s], rho[1][s]);
#pragma unroll
                for (int r = 0; r < K; ++r) ch.A[r][s] = mk2(A[0][r][s], A[1][r][s]);
            }
            FwdOut fo = ch.ragged ? forward_pass<true, false>(ch, em, rv) : forward_pass<false, false>(ch, em, rv);
            {
                f2 chk = fo.pf.v[0];
#pragma unroll
                for (int s = 1; s < K; ++s) chk = chk + fo.pf.v[s];
                const bool okx = chk.v.x > 0.5f && chk.v.x < 2.f, oky = chk.v.y > 0.5f && chk.v.y < 2.f;
                if (__builtin_expect(T > 0 && !(okx && oky), 0)) {
                    fo = forward_pass<true, true>(ch, em, rv);
                    events[0] += fo.events;            // attributed to the first chain of the pair (count only)
                }
            }

            // ---- 3. relabel and emit, per chain; forecasts packed
#pragma unroll
            for (int c = 0; c < 2; ++c) ranks_of<float, K>(mu[c], ch.rank[c]);
            const long long draw_idx = gs - a.burnin;
            const bool save = (draw_idx >= 0) && (T > 0);
            if (save) {
                const size_t i = (size_t)(draw_idx - a.draw0);
                const size_t cs = (size_t)a.chunk * ns;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float* o = out + i * ns + slot0 + c;
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        o[(size_t)(ch.rank[c][s]) * cs] = mu[c][s];
                        o[(size_t)(K + ch.rank[c][s]) * cs] = sig2[c][s];
                        o[(size_t)(2 * K + K * K + ch.rank[c][s]) * cs] = c == 0 ? fo.pf.v[s].v.x : fo.pf.v[s].v.y;
#pragma unroll
                        for (int r = 0; r < K; ++r) o[(size_t)(2 * K + ch.rank[c][s] * K + ch.rank[c][r]) * cs] = A[c][r][s];
                    }
                }
                const int f0 = 3 * K + K * K;
                f2 v[K], mu2[K];
#pragma unroll
                for (int s = 0; s < K; ++s) { v[s] = fo.pf.v[s]; mu2[s] = mk2(mu[0][s], mu[1][s]); }
                int h = 0;
                float* o = out + i * ns + slot0;
                for (int j = 0; j < a.n_h; ++j) {
                    for (; h < a.h_sorted[j]; ++h) {
                        f2 nv[K];
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            f2 acc = v[0] * ch.A[0][s];
#pragma unroll
                            for (int r = 1; r < K; ++r) acc = fma2(v[r], ch.A[r][s], acc);
                            nv[s] = acc;
                        }
#pragma unroll
                        for (int s = 0; s < K; ++s) v[s] = nv[s];
                    }
                    f2 f = v[0] * mu2[0];
#pragma unroll
                    for (int s = 1; s < K; ++s) f = fma2(v[s], mu2[s], f);
                    const float yr = reinterpret_cast<const float*>(a.yfut)[(size_t)a.h_slot[j] * ns + slot0];
                    float* of = o + (size_t)(f0 + 2 * a.h_slot[j]) * cs;
                    of[0] = f.v.x; of[1] = f.v.y;
                    of[cs] = f.v.x - yr; of[cs + 1] = f.v.y - yr;
                }
                if (LOGLIK) { o[(size_t)(f0 + 2 * a.n_h) * cs] = fo.ll[0]; o[(size_t)(f0 + 2 * a.n_h) * cs + 1] = fo.ll[1]; }
            }

            // ---- 4. backward pass
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int x = 0; x < K; ++x) {
                    Entry en;
#pragma unroll
                    for (int r = 0; r < K; ++r) en.a[r] = A[c][r][x];
                    en.inc = (Row)1 << (Pack::kBits * x);
                    smem_tab[(c * K + x) * kGibbsThreads + threadIdx.x] = en;
                }
            BackOut bo = ch.ragged ? backward_pass<true, false>(ch, fo.pf, key[0], key[1], sweep, a.flags)
                                   : backward_pass<false, false>(ch, fo.pf, key[0], key[1], sweep, a.flags);
            if (__builtin_expect(__any_sync(0xffffffffu, bo.bad), 0)) bo = backward_pass<true, true>(ch, fo.pf, key[0], key[1], sweep, a.flags);   // warp-uniform (shared ring)

            // ---- unpack the statistics for the next sweep's draws
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const Back& b = bo.b[c];
                float sS = 0.f, sQ = 0.f;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    int n = (T > 0 && bo.xN[c] == i) ? 1 : 0;
#pragma unroll
                    for (int j = 0; j < K; ++j) { trans[c][i][j] = b.tr.get(i, j); n += trans[c][i][j]; }
                    cnt[c][i] = n;
                }
#pragma unroll
                for (int i = 0; i < K - 1; ++i) { Sd[c][i] = b.Sd[i]; Qd[c][i] = b.Qd[i]; sS += b.Sd[i]; sQ += b.Qd[i]; }
                Sd[c][K - 1] = totS - sS; Qd[c][K - 1] = totQ - sQ;
                if (cnt[c][K - 1] == 0) { Sd[c][K - 1] = 0.f; Qd[c][K - 1] = 0.f; }
            }
        }

#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int slot = slot0 + c;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                a.cnt[i * ns + slot] = cnt[c][i];
                reinterpret_cast<float*>(a.Sd)[i * ns + slot] = Sd[c][i];
                reinterpret_cast<float*>(a.Qd)[i * ns + slot] = Qd[c][i];
#pragma unroll
                for (int j = 0; j < K; ++j) a.trans[(i * K + j) * ns + slot] = trans[c][i][j];
            }
            a.events[slot] += events[c];
        }
    }
};

template <int K, bool WIDE> constexpr size_t kPairSmemBytes() {
    return sizeof(typename GibbsPair<K, false, WIDE>::Entry) * 2 * K * kGibbsThreads
           + sizeof(float2) * (size_t)(kGibbsThreads / 32) * kRing * 4 * K * 32;
}

#ifndef HMC_PAIR_MINBLOCKS
#define HMC_PAIR_MINBLOCKS 3
#endif

template <int K, bool LOGLIK, bool WIDE>
__global__ void __launch_bounds__(kGibbsThreads, HMC_PAIR_MINBLOCKS) gibbs_pair_kernel(const GibbsArgs a) {
    using W = GibbsPair<K, LOGLIK, WIDE>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename W::Entry* tab = reinterpret_cast<typename W::Entry*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (kGibbsThreads / 32) + (threadIdx.x >> 5);
    if (g >= a.n_tasks) return;
    W::run(a, a.task0 + g * a.task_stride, lane, tab);
}

// host-side launcher (gibbs_inst.cu, float units only)
template <int K> cudaError_t launch_gibbs_pair(const GibbsLaunch& cfg, const GibbsArgs& a, cudaStream_t st);

}  // namespace hmc
