/*
 * hmcgpu.h — C ABI of the B200 (sm_100a) Gibbs/FFBS path that replaces the hot loop of joe5saia/Hmc.jl.
 *
 * Plain C: opaque handles, plain pointers and sizes, int return codes.  No C++ types, no exceptions,
 * no stdout, no exit() cross this boundary.  Julia binds it with `ccall` (hmc.jl_b200/julia/HmcGPU.jl),
 * the test harness with Python ctypes (hmc.jl_b200/binding.py); both see exactly these symbols.
 *
 * Reference interfaces replaced (paths under the reference checkout, src/Hmc.jl):
 *   hmcgpu_estimate / hmcgpu_plan_*  <- estimatemodel (:850-865) = makeParams (:161-195) + HyperParams (:132-142)
 *                                       + gibbssample! (:517-562, gibbssweep! :486-515) + forecast loop (:858-862);
 *                                       called from code/run_hmm.jl:119
 *   hmcgpu_filter / _filter_masked   <- forwardupdate_P! (:371-440)   (fixed parameters -> pif, totals, loglik; signal rows)
 *   hmcgpu_smooth                    <- backwardupdate_P! (:442-457)  (marginals pib)
 *   hmcgpu_sample_states             <- update_X! (:459-484)          (injected uniforms -> state path)
 *   hmcgpu_draw_params / _signals    <- update_μσ! draws (:302-335; with signal statistics :267-314), update_ρ! (:350-356),
 *                                       update_A! (:358-369)
 *   hmcgpu_forecast                  <- forecast (:658-667)
 *
 * Conventions
 *   - All host arrays are fp64 (int64 for states) whatever the device precision.
 *   - Batched deterministic calls use row-major [batch][...] arrays: A[b][r][s] = P(X_t=s|X_{t-1}=r),
 *     pif[b][t][s], sigma2 holds VARIANCES (as the reference's σ does).
 *   - hmcgpu_estimate outputs use Julia COLUMN-MAJOR layout per window so Julia can wrap them without
 *     a copy (see hmcgpu_result).
 *   - States are 1-based in X, windows are 1-based inclusive [start,end] like `sampleRange`.
 *   - Ownership: the caller owns every pointer it passes; the library copies what it needs before
 *     returning and never frees or retains host pointers.  Device memory belongs to ctx / plan.
 *   - Threading: a ctx is bound to one GPU and is not re-entrant; different ctxs may be driven from
 *     different host threads concurrently.
 *   - Return codes: 0 = OK; < 0 = error (message via hmcgpu_last_error); > 0 (estimate/plan_fetch only)
 *     = number of chains that saw a zero-normaliser / non-finite event (data still returned — mirrors the
 *     reference's warn-and-continue, :319-329, :435, :472-480).
 *   - There is no CPU fallback: without a CUDA device every compute entry point returns HMCGPU_ERR_NODEVICE.
 */
#ifndef HMCGPU_H
#define HMCGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMCGPU_OK 0
#define HMCGPU_ERR_ARG (-1)
#define HMCGPU_ERR_CUDA (-2)
#define HMCGPU_ERR_ALLOC (-3)
#define HMCGPU_ERR_UNSUPPORTED (-4)
#define HMCGPU_ERR_NODEVICE (-5)

/* hmcgpu_problem.flags */
#define HMCGPU_FLAG_REF_Q1 1u        /* reproduce :512-514 (X[N] drawn from the relabelled pif[N,:]); hosts set it by default */
#define HMCGPU_FLAG_DRAWS 2u         /* fill the per-draw arrays of hmcgpu_result */
#define HMCGPU_FLAG_SUMMARY 4u       /* fill per-window posterior mean / variance (pooled over chains and draws) */
#define HMCGPU_FLAG_SMOOTHED_MEAN 8u /* fill pib_mean: posterior mean of the smoothed state probabilities */
#define HMCGPU_FLAG_LOGLIK 16u       /* compute the per-draw log-likelihood sum_t log(total_t) */
#define HMCGPU_FLAG_FILTERED_MEAN 64u /* fill pib_mean / insample_forecast_mean with the posterior means of the FILTERED state
                                        probabilities pif[t,:] and of pif[t,:]' A^h mu (what the reference's published in-sample table
                                        data/output/official_insample/forecats_insample.csv holds); K <= 8, not with SMOOTHED_MEAN or signals */

typedef struct hmcgpu_ctx hmcgpu_ctx;
typedef struct hmcgpu_plan hmcgpu_plan;

/* One rolling-window estimation job = what code/run_hmm.jl builds in `opt` (:95-109), for many end dates at once. */
typedef struct hmcgpu_problem {
    const double* y;          /* series, COLUMN-major [y_len x n_series] (series s at y + s*y_len) = opt.rawdata */
    int64_t y_len;
    int32_t n_series;         /* >= 1 */
    int32_t n_windows;        /* independent estimation windows (end dates) */
    const int32_t* win_series;/* [n_windows] 0-based series index, NULL = all 0 */
    const int32_t* win_start; /* [n_windows] 1-based first index of sampleRange */
    const int32_t* win_end;   /* [n_windows] 1-based last index of sampleRange (= endIndex) */
    const int64_t* win_id;    /* [n_windows] global window id used in the RNG counter, NULL = 0..n_windows-1.
                                 Makes results independent of how windows are sharded over GPUs. */
    int32_t K;                /* number of states D */
    int32_t n_chains;         /* independent chains per window (reference: 1) */
    int64_t burnin, nrun;     /* opt.burnin, opt.Nrun */
    uint64_t seed;            /* opt.seed -> Philox key */
    const double* xi;         /* [K] prior mean; NULL = mean(Y_window) (:136) */
    const double* alpha;      /* [K] NULL = 1 (:137) */
    const double* nu;         /* [K] NULL = 1 (:140) */
    const double* beta0;      /* [K] β used by the first sweep; NULL = 1 (:179) */
    const double* beta;       /* [K] β afterwards; NULL = 2 (:347) */
    double kappa;             /* hp.κ = opt.noise (:116, :158): signals are emitted with sd*(1+κ) (:382) and weigh 1/(1+κ) in
                                 the statistics (:302, :314); only used with is_signal */
    const uint8_t* is_signal; /* [y_len] 1 = time index belongs to opt.signalRange (noisy signal), 0 = observation: one mask for
                                 every series, or [y_len x n_series] column-major like y when is_signal_per_series != 0.
                                 NULL = no signals (estimatemodel).  K <= 4, not with SMOOTHED_MEAN. */
    const int32_t* horizons;  /* [n_h] forecast horizons >= 0 from the window end (reference scripts: {12}); 0 = pi_end'mu */
    int32_t n_h;
    const int64_t* X0;        /* [sum_w T_w] initial state paths (1-based states), windows concatenated in caller order and
                                 shared by the window's chains; NULL = makeParams rule (:185-187) */
    int32_t precision;        /* 32 or 64: device arithmetic */
    uint32_t flags;
    /* ---- signals tier (estimatesignals!, :868-914); zero / NULL = estimatemodel behaviour */
    const int32_t* win_init_series; /* [n_windows] series the makeParams / HyperParams rules read (X0, default xi), while the
                                 chain itself runs on win_series: estimatesignals! initialises from the real data and
                                 estimates on the perturbed copy (:888-892).  NULL = win_series */
    int32_t pi_row_back;      /* pi_end (and its summary) = smoothed marginal pib[N - pi_row_back, :] instead of the last row:
                                 samples.πb[:, opt.endIndex, :] with signalLen rows after endIndex (:893).  K <= 4. */
    int32_t is_signal_per_series; /* 0: is_signal holds y_len flags; 1: y_len flags per series (many end dates, each with its
                                 own signalRange, in one call) */
} hmcgpu_problem;

/* Caller-allocated outputs; any pointer may be NULL.  R = n_chains*nrun draws per window, draw index
 * d = chain*nrun + i (chain-major).  Per window w the arrays are Julia column-major, windows concatenated:
 *   mu[w*R*K + k*R + d], sigma2 likewise, pi_end likewise (= samples.πb[:,end,:], :744)
 *   A[w*R*K*K + (s*K + r)*R + d]            (reshape(A,R,:) gives trans columns as saveresults expects, :745)
 *   forecasts[w*R*2*n_h + (2*j+e)*R + d]    e=0 forecast, e=1 forecast error for horizons[j] (:858-862)
 *   loglik[w*R + d]
 * States are emitted in increasing-μ order (:501-513).
 * Summary fields per window, F = 3K + K*K + 2*n_h + 1 in the order
 *   mu[K], sigma2[K], A[K*K] (index s*K+r), pi_end[K], forecasts[2*n_h], loglik:
 *   summary_mean[w*F + f], summary_var[w*F + f]  (population variance over the R pooled draws; the loglik field is 0
 *   unless HMCGPU_FLAG_LOGLIK is set)
 * pib_mean: per window N_w x K column-major, concatenated: offset_w = K * sum_{v<w} N_v.
 * insample_forecast_mean: per window N_w x n_h column-major, concatenated (offset_w = n_h * sum_{v<w} N_v): the posterior
 *   mean of pib[t,:]' A^h mu for every date t of the window and every horizon (forecastinsample, src/Hmc.jl:683-699);
 *   the forecast error of date t is this minus y[t+h].  Needs HMCGPU_FLAG_SMOOTHED_MEAN (K <= 4).
 * status[w*n_chains + c]: event count of that chain. */
typedef struct hmcgpu_result {
    double* mu;
    double* sigma2;
    double* A;
    double* pi_end;
    double* forecasts;
    double* loglik;
    double* summary_mean;
    double* summary_var;
    double* pib_mean;
    double* insample_forecast_mean;
    int32_t* status;
    /* filled by the library */
    double gpu_ms;        /* device time of the whole run: chain init, every sweep launch, per-chunk post-processing (CUDA
                             events on the library's stream) */
    double sweep_kernel_ms; /* device time from the start of the first Gibbs sweep launch to the end of the last one (CUDA events
                             on the task-group streams; the launches of different groups overlap, so this is their enclosing
                             interval — it leaves out the init kernel and the post-processing after the last sweep) */
    int64_t n_launches;   /* kernels launched for this job */
    int64_t n_sweep_launches; /* launches of the Gibbs sweep kernel among them */
    int64_t h2d_bytes, d2h_bytes;
    int64_t state_steps;  /* sum_w T_w * n_chains * (burnin + nrun) */
    /* ---- appended in ABI version 200 (older callers that zero-initialise the struct keep working) */
    double sweep_launch_ms_sum; /* sum over the sweep-kernel launches of each launch's own duration (an event pair around every
                             launch on its stream); divided by n_sweep_launches = the average launch duration */
    int32_t sweep_kernel; /* which sweep kernel the plan chose: HMCGPU_KERNEL_* */
    int32_t n_tasks;      /* warp tasks of that kernel (32 lanes each) */
} hmcgpu_result;

#define HMCGPU_KERNEL_THREAD 0  /* one thread per chain (gibbs_sweeps_kernel), wide batches, K <= 8 */
#define HMCGPU_KERNEL_SCAN 1    /* one warp per chain, time-parallel scans (gibbs_scan_kernel), narrow batches */
#define HMCGPU_KERNEL_LANE 2    /* one lane per state (gibbs_wide_kernel), K = 9..32 */
#define HMCGPU_KERNEL_PAIR 3    /* reserved (the two-chains-per-thread kernel of round 1 was removed: slower than the default at every width) */
#define HMCGPU_KERNEL_SEG 4     /* L lanes per chain, each lane a contiguous time segment (gibbs_seg_kernel), mid-width batches */

int hmcgpu_version(void);
/* "<abi version>;<hash of the CUDA/C sources the library was built from>" — hosts compare the hash with the sources next to
 * them and refuse (or rebuild) a stale library */
const char* hmcgpu_build_info(void);
/* number of CUDA devices visible (0 if none / no driver) */
int hmcgpu_device_count(void);
int hmcgpu_ctx_create(int device, hmcgpu_ctx** out);
void hmcgpu_ctx_destroy(hmcgpu_ctx* ctx);
/* message of the last error on this ctx (ctx may be NULL: last error of ctx_create on this thread) */
const char* hmcgpu_last_error(const hmcgpu_ctx* ctx);
/* blocks until all work queued by this ctx has finished */
int hmcgpu_ctx_sync(hmcgpu_ctx* ctx);
/* Releases the idle device buffers the context keeps for reuse between calls (bounded by HMCGPU_POOL_MAX_GB, default 16 GiB;
 * they are also released when an allocation on this context fails and at hmcgpu_ctx_destroy). */
int hmcgpu_ctx_trim(hmcgpu_ctx* ctx);

/* Whole estimation, host buffers in and out (H2D, all sweeps, D2H inside the call). */
int hmcgpu_estimate(hmcgpu_ctx* ctx, const hmcgpu_problem* p, hmcgpu_result* r);
/* Same, windows sharded over several GPUs by sum of T_w (one host thread per device, no collectives). */
int hmcgpu_estimate_multi(const int* devices, int n_devices, const hmcgpu_problem* p, hmcgpu_result* r);

/* Split form: create uploads inputs and allocates device state; run executes every sweep (inputs resident
 * in HBM, no host transfers); fetch copies the requested outputs back.  run may be repeated (it restarts
 * the chains from the initial state). */
int hmcgpu_plan_create(hmcgpu_ctx* ctx, const hmcgpu_problem* p, hmcgpu_plan** out);
int hmcgpu_plan_run(hmcgpu_plan* plan);
int hmcgpu_plan_fetch(hmcgpu_plan* plan, hmcgpu_result* r);
void hmcgpu_plan_destroy(hmcgpu_plan* plan);

/* Forward filter for fixed parameters.  y: [T] shared by the batch (y_batch_stride = 0) or [B][T]
 * (y_batch_stride = T).  totals [B][T] and loglik [B] may be NULL. */
int hmcgpu_filter(hmcgpu_ctx* ctx, int32_t precision, int32_t K, int64_t B, int64_t T,
                  const double* y, int64_t y_batch_stride,
                  const double* A, const double* mu, const double* sigma2, const double* rho,
                  double* pif, double* totals, double* loglik);
/* Same with the signal branch of forwardupdate_P! (:380-383, :396, :424): is_signal [T] flags the rows emitted with
 * sd*(1+kappa) (one mask for the whole batch; NULL = hmcgpu_filter).  K <= 4. */
int hmcgpu_filter_masked(hmcgpu_ctx* ctx, int32_t precision, int32_t K, int64_t B, int64_t T,
                         const double* y, int64_t y_batch_stride, const uint8_t* is_signal, double kappa,
                         const double* A, const double* mu, const double* sigma2, const double* rho,
                         double* pif, double* totals, double* loglik);
/* Smoothed marginals pib [B][T][K] from pif [B][T][K] and A [B][K][K]. */
int hmcgpu_smooth(hmcgpu_ctx* ctx, int32_t precision, int32_t K, int64_t B, int64_t T,
                  const double* A, const double* pif, double* pib);
/* Backward state sampling with injected uniforms u [B][T] in [0,1); fp64, bit-exact against the oracle's
 * pif form.  piN [B][K] (NULL = pif[:,T-1,:]) is what X[T] is drawn from.  X [B][T], 1-based. */
int hmcgpu_sample_states(hmcgpu_ctx* ctx, int32_t K, int64_t B, int64_t T,
                         const double* A, const double* pif, const double* piN, const double* u, int64_t* X);
/* Conjugate draws from sufficient statistics with the library's Philox streams (chain ids chain0..chain0+B-1).
 * Ni [B][K] counts, S [B][K] sums, S2 [B][K] centred sums of squares, trans [B][K][K] counts INCLUDING the +1 prior. */
int hmcgpu_draw_params(hmcgpu_ctx* ctx, int32_t precision, int32_t K, int64_t B,
                       const int64_t* Ni, const double* S, const double* S2, const int64_t* trans,
                       const double* xi, const double* alpha, const double* nu, const double* beta,
                       uint64_t seed, uint32_t chain0, uint32_t sweep,
                       double* sigma2, double* mu, double* rho, double* A);
/* Same with the noisy signals of update_μσ! (:267-314): Mi [B][K] signal counts, Sm [B][K] sums, Sm2 [B][K] centred sums of
 * squares of the signals of each state, kappa their relative imprecision (Ni / S / S2 then cover the observations only).  K <= 4. */
int hmcgpu_draw_params_signals(hmcgpu_ctx* ctx, int32_t precision, int32_t K, int64_t B,
                               const int64_t* Ni, const double* S, const double* S2,
                               const int64_t* Mi, const double* Sm, const double* Sm2, double kappa, const int64_t* trans,
                               const double* xi, const double* alpha, const double* nu, const double* beta,
                               uint64_t seed, uint32_t chain0, uint32_t sweep,
                               double* sigma2, double* mu, double* rho, double* A);
/* out [B][2*n_h]: forecast and error for each horizon; yreal [n_h]. */
int hmcgpu_forecast(hmcgpu_ctx* ctx, int32_t K, int64_t B, const double* mu, const double* A, const double* pi,
                    const int32_t* horizons, int32_t n_h, const double* yreal, double* out);
/* Raw Philox4x32-10 blocks (known-answer tests): ctr [n][4], key [n][2] -> out [n][4]. */
int hmcgpu_philox(hmcgpu_ctx* ctx, int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out);
/* Same with the round count of the stream: 10 = parameter draws, 7 = the backward sampler's state uniforms. */
int hmcgpu_philox_rounds(hmcgpu_ctx* ctx, int32_t rounds, int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
