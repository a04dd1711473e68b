// hmm_device.cuh — per-chain device primitives of the Gibbs/FFBS path (one thread = one chain, K in registers).
// Every kernel in hmcgpu.cu (fused Gibbs sweeps and the deterministic entry points) is built from these.
// Reference lines (src/Hmc.jl) are cited per function.
#pragma once
#include "rng.cuh"

namespace hmc {

// Blackwell packed two-wide fp32 arithmetic (FFMA2 / FMUL2 / FADD2: one issue slot, two results)
struct f2 { float2 v; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ f2 splat2(float a) { return mk2(a, a); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }

template <typename R> struct Real;
template <> struct Real<float> {
    static __device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float eps() { return 2.220446049250313e-16f; }
};
template <> struct Real<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }
};

// ---------------------------------------------------------------- emission model (src/Hmc.jl:380-383, pdf at :390/:415)
// fp32: log2-domain quadratic  l_s = (y-mu_s)^2 * q_s + c_s, max-subtracted before ex2 (scaled recursion, hazard H1).
// fp64: the reference's direct pdf  exp(-z^2/2) * invsqrt2pi / sd  (no rescaling, like the reference).
template <typename R, int K> struct Emission;

template <int K> struct Emission<float, K> {
    float mu[K], q[K], c[K];
    __device__ __forceinline__ void prepare(const float (&m)[K], const float (&sig2)[K]) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            mu[s] = m[s];
            q[s] = -0.72134752044448170368f / sig2[s];                       // -0.5*log2(e)/sigma^2
            c[s] = -0.5f * Real<float>::lg2(6.283185307179586f * sig2[s]);   // log2(1/(sd*sqrt(2pi)))
        }
    }
    // e[s] proportional to pdf_s(y); returns log2 of the common factor that was divided out
    template <bool EXACT = false> __device__ __forceinline__ float eval(float y, float (&e)[K]) const {
        float l[K];
#pragma unroll
        for (int s = 0; s < K; ++s) { const float d = y - mu[s]; l[s] = fmaf(d * d, q[s], c[s]); }
        float m = l[0];
#pragma unroll
        for (int s = 1; s < K; ++s) m = fmaxf(m, l[s]);
#pragma unroll
        for (int s = 0; s < K; ++s) e[s] = Real<float>::ex2(l[s] - m);
        return m;
    }
};

// exp(x) for x <= 0 in fp64 without the libm call: x = n ln2 + f, |f| <= ln2/2, degree-11 Taylor polynomial of exp(f) (truncation
// 2e-14 relative, result within ~1e-15 relative over [-708, 0]), 2^n added into the exponent field.  The fp64 sweeps spend most
// of their arithmetic in the K exponentials per time step; north_star asks 1e-5 of the fp64 path.  The deterministic entry
// points (hmcgpu_filter / smooth ...) keep libm's exp (Emission::eval<true>).
#ifndef HMC_FAST_EXP64
#define HMC_FAST_EXP64 1
#endif
__device__ __forceinline__ double exp_nonpos_fast(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);         // round(x log2 e) in the low mantissa bits
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;
    double f = fma(nd, -6.93147180369123816490e-01, x);                       // ln2 hi
    f = fma(nd, -1.90821492927058770002e-10, f);                              // ln2 lo
    double p = 2.505210838544172e-08;                                          // 1/11!
    p = fma(p, f, 2.755731922398589e-07);
    p = fma(p, f, 2.755731922398589e-06);
    p = fma(p, f, 2.48015873015873e-05);
    p = fma(p, f, 1.984126984126984e-04);
    p = fma(p, f, 1.388888888888889e-03);
    p = fma(p, f, 8.333333333333333e-03);
    p = fma(p, f, 4.166666666666666e-02);
    p = fma(p, f, 1.666666666666667e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const int hi = __double2hiint(p) + (n << 20);                              // p in [0.7, 1.42): scale by 2^n (n >= -1021 here)
    const double r = __hiloint2double(hi, __double2loint(p));
    return x < -707.0 ? 0.0 : r;
}
template <bool EXACT> __device__ __forceinline__ double exp_nonpos(double x) {
#if HMC_FAST_EXP64
    if constexpr (!EXACT) return exp_nonpos_fast(x);
#endif
    return exp(x);
}

template <int K> struct Emission<double, K> {
    double mu[K], isd[K], nrm[K];
    __device__ __forceinline__ void prepare(const double (&m)[K], const double (&sig2)[K]) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            mu[s] = m[s];
            const double sd = sqrt(sig2[s]);
            isd[s] = 1.0 / sd;
            nrm[s] = 0.3989422804014327 / sd;
        }
    }
    template <bool EXACT = false> __device__ __forceinline__ double eval(double y, double (&e)[K]) const {
#pragma unroll
        for (int s = 0; s < K; ++s) { const double z = (y - mu[s]) * isd[s]; e[s] = exp_nonpos<EXACT>(-0.5 * (z * z)) * nrm[s]; }
        return 0.0;
    }
    // signal rows: Normal(mu, (1+kappa) sd) (src/Hmc.jl:382); sw = +-1/(1+kappa) or 1 (the sign only flags a signal)
    template <bool EXACT = false> __device__ __forceinline__ void eval_scaled(double y, double sw, double (&e)[K]) const {
        const double a = fabs(sw);
#pragma unroll
        for (int s = 0; s < K; ++s) { const double z = (y - mu[s]) * isd[s] * a; e[s] = exp_nonpos<EXACT>(-0.5 * (z * z)) * (nrm[s] * a); }
    }
};

// ---------------------------------------------------------------- forward step (src/Hmc.jl:386-432)
// pi <- normalise_s( (sum_r pi[r] A[r][s]) * e[s] ).  Returns the normaliser `total`; ok=false on a zero /
// non-finite normaliser (the reference only warns, :435) in which case pi is reset to uniform.
template <typename R, int K>
__device__ __forceinline__ R forward_step(const R (&A)[K][K], const R (&e)[K], R (&pi)[K], bool& ok) {
    R q[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
        R pred = pi[0] * A[0][s];
#pragma unroll
        for (int r = 1; r < K; ++r) pred = fma(pi[r], A[r][s], pred);
        q[s] = pred * e[s];
    }
    R tot = q[0];
#pragma unroll
    for (int s = 1; s < K; ++s) tot += q[s];
    ok = (tot > R(0)) && (tot < R(3.0e38));
    const R inv = Real<R>::rcp(tot);
#pragma unroll
    for (int s = 0; s < K; ++s) pi[s] = q[s] * inv;
    return tot;
}

// ---------------------------------------------------------------- backward smoothing step (src/Hmc.jl:449-455)
// Given pif_t (pf), A and pib_{t+1} (pb, updated in place to pib_t):
//   pib_t[r] = pf[r] * sum_s A[r][s] * pib_{t+1}[s] / pred[s],  pred[s] = sum_r pf[r] A[r][s]
// (algebraically the reference's Pb recursion; needs no emission and no stored Pf).
template <typename R, int K>
__device__ __forceinline__ void smooth_step(const R (&A)[K][K], const R (&pf)[K], R (&pb)[K]) {
    R w[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
        R pred = pf[0] * A[0][s];
#pragma unroll
        for (int r = 1; r < K; ++r) pred = fma(pf[r], A[r][s], pred);
        w[s] = pred > R(0) ? pb[s] * Real<R>::rcp(pred) : R(0);
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
        R acc = A[r][0] * w[0];
#pragma unroll
        for (int s = 1; s < K; ++s) acc = fma(A[r][s], w[s], acc);
        pb[r] = pf[r] * acc;
    }
}

// ---------------------------------------------------------------- categorical draws (src/Hmc.jl:464, :468-481)
// Distributions 0.21 rand(Categorical(p)): i=1; c=p[1]; while c < u && i < n: c += p[i+=1].
// Unnormalised form: compare the running sum with u*total  (0-based result).
template <typename R, int K>
__device__ __forceinline__ int categorical_unnorm(const R (&p)[K], R u) {
    R tot = p[0];
#pragma unroll
    for (int i = 1; i < K; ++i) tot += p[i];
    const R thr = u * tot;
    int x = 0;
    R c = p[0];
#pragma unroll
    for (int i = 1; i < K; ++i) { x += (c < thr) ? 1 : 0; c += p[i]; }
    return x;
}

// Exact fp64 form used by hmcgpu_sample_states: same operations, same order and roundings as the oracle
// (orc_sample_states form 1): p[r] = pif[k,r]*A[r,x]; total; p/total; cumulative compare against u.
template <int K>
__device__ __forceinline__ int categorical_exact(const double (&p)[K], double u) {
    int i = 0;
    double c = p[0];
#pragma unroll
    for (int j = 1; j < K; ++j) {
        if (c < u && i == j - 1) { i = j; c = __dadd_rn(c, p[j]); }
    }
    return i;
}

template <typename R, int K>
__device__ __forceinline__ R select_k(const R (&v)[K], int x) {
    R r = v[K - 1];
#pragma unroll
    for (int i = K - 2; i >= 0; --i) r = (x == i) ? v[i] : r;
    return r;
}

// One backward sampling step (src/Hmc.jl:466-481 in the pif form): draws X_k given X_{k+1}=xn.
// gate = pif[k+1, xn] (the reference's `total`, quirk Q5): uniform p when gate <= eps().
template <typename R, int K>
__device__ __forceinline__ int backward_sample_step(const R (&Acol)[K], const R (&pf)[K], R gate, R u) {
    R p[K];
#pragma unroll
    for (int r = 0; r < K; ++r) p[r] = pf[r] * Acol[r];
    if (!(gate > Real<R>::eps())) {
#pragma unroll
        for (int r = 0; r < K; ++r) p[r] = R(1);
    }
    return categorical_unnorm<R, K>(p, u);
}

// ---------------------------------------------------------------- conjugate draws (src/Hmc.jl:302-335, :350-369)
// Sufficient statistics are kept about a shift c:  Sd = sum(y-c), Qd = sum((y-c)^2) over t with X_t = i.
template <typename R, int K> struct Hyper { R xi[K], alpha[K], nu[K], beta[K]; };
// statistics of the noisy signals of a window (src/Hmc.jl:267-300): counts Mi, shifted sums over the signal time steps;
// k1 = 1/(1+kappa) is their relative precision weight (:302, :314)
template <typename R, int K> struct SigStats { int m[K]; R Sm[K], Qm[K]; R k1; };

// cnt / Sd / Qd describe the plain observations; with SIG the signals enter through sg (cnt then excludes them)
template <typename R, int K, bool SIG = false>
__device__ __forceinline__ int draw_params(const int (&cnt)[K], const R (&Sd)[K], const R (&Qd)[K], const int (&trans)[K][K],
                                           R c, const Hyper<R, K>& hp, const RngKey& key, uint32_t sweep,
                                           R (&sig2)[K], R (&mu)[K], R (&rho)[K], R (&A)[K][K],
                                           const SigStats<R, K>* sg = nullptr) {
    int events = 0;
    R neff[K], sumd[K], ntot[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const R n = (R)cnt[i];
        const R dbar = cnt[i] > 0 ? Sd[i] / n : R(0);               // ybar - c
        R s2 = Qd[i] - n * dbar * dbar;                              // sum (y - ybar)^2   (:291-294)
        s2 = s2 > R(0) ? s2 : R(0);
        R totalbar = cnt[i] > 0 ? dbar + c : R(0);                   // :282-288
        R a = hp.alpha[i] + R(0.5) * n;                              // :313
        R ne = n, extra = R(0);
        sumd[i] = Sd[i]; ntot[i] = n;
        if constexpr (SIG) {
            const R m = (R)sg->m[i];
            const R sbar = sg->m[i] > 0 ? sg->Sm[i] / m : R(0);      // sbar - c   (:272-278)
            R sm2 = sg->Qm[i] - m * sbar * sbar;                     // sum (signal - sbar)^2   (:297-300)
            sm2 = sm2 > R(0) ? sm2 : R(0);
            sumd[i] = Sd[i] + sg->Sm[i]; ntot[i] = n + m;
            totalbar = (cnt[i] + sg->m[i]) > 0 ? sumd[i] / ntot[i] + c : R(0);
            ne = n + m * sg->k1;                                     // Neff = Ni + Mi/(1+kappa)  (:302-303)
            a += R(0.5) * m;
            extra = R(0.5) * sg->k1 * sm2;                           // (0.5/(1+kappa)) Sm2  (:314)
        }
        const R dev = totalbar - hp.xi[i];
        R b = hp.beta[i] + R(0.5) * s2;                              // :314
        if constexpr (SIG) b += extra;
        b += R(0.5) * ne * hp.nu[i] / (ne + hp.nu[i]) * (dev * dev);
        neff[i] = ne;
        if (a > R(0) && b > R(0)) {
            const R g = gamma_mt<R>(a, key, sweep, (KIND_SIGMA << 16) | (uint32_t)i);
            sig2[i] = b / g;                                         // :320 InverseGamma(a,b)
        } else {
            ++events;                                                // reference: catch + keep old value (:321-329)
        }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const R n = neff[i];
        const R sum_y = sumd[i] + ntot[i] * c;                       // S + Sm: the plain sum over observations and signals (:331)
        const R m = (sum_y + hp.nu[i] * hp.xi[i]) / (n + hp.nu[i]);  // :331
        const R s = M<R>::sqrt(sig2[i] / (n + hp.nu[i]));            // :332
        const uint4 w = rng_block(key, sweep, (KIND_MU << 16), (uint32_t)(i >> 1));
        const R z = (i & 1) ? normal_from<R>(w.z, w.w) : normal_from<R>(w.x, w.y);
        mu[i] = m + s * z;                                           // :334
    }
    {                                                                // :354-355 Dirichlet(1,...,1)
        R tot = R(0);
#pragma unroll
        for (int i = 0; i < K; ++i) { rho[i] = gamma_mt<R>(R(1), key, sweep, (KIND_RHO << 16) | (uint32_t)i); tot += rho[i]; }
#pragma unroll
        for (int i = 0; i < K; ++i) rho[i] /= tot;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {                                    // :362-368 Dirichlet(1 + counts)
        R tot = R(0);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            A[i][j] = gamma_mt<R>((R)(trans[i][j] + 1), key, sweep, (KIND_A << 16) | (uint32_t)(i * K + j));
            tot += A[i][j];
        }
#pragma unroll
        for (int j = 0; j < K; ++j) A[i][j] /= tot;
    }
    return events;
}

// rank[i] = position of state i in increasing-mu order (stable), i.e. the inverse of sortperm(mu) (src/Hmc.jl:501)
template <typename R, int K>
__device__ __forceinline__ void ranks_of(const R (&mu)[K], int (&rank)[K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) r += (mu[j] < mu[i] || (mu[j] == mu[i] && j < i)) ? 1 : 0;
        rank[i] = r;
    }
}

}  // namespace hmc
