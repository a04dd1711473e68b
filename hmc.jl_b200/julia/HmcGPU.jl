# HmcGPU.jl — Julia binding of libhmcgpu.so (include/hmcgpu.h): the B200 replacement of the Gibbs/FFBS hot loop of Hmc.jl.
#
# Drop-in use from code/run_hmm.jl (replaces `samples = Hmc.estimatemodel(opt)` at line 119):
#
#     include("HmcGPU.jl")
#     samples = HmcGPU.estimatemodel(opt)            # same NamedTuple fields: μ, σ, πb, A, forecasts, obsdates
#     Hmc.saveresults(samples, opt, p; hassignals = false)
#
# NOTE: Julia is not installed in the build environment of this repository, so this file has never been executed
# there.  It binds exactly the C ABI that tests/ exercise through Python ctypes (hmc.jl_b200/binding.py); the struct
# layouts below mirror include/hmcgpu.h field by field (checked for the ctypes mirror in tests/test_host.py).
module HmcGPU

using Dates

const LIB = get(ENV, "HMCGPU_LIB", joinpath(@__DIR__, "..", "lib", "libhmcgpu.so"))

const FLAG_REF_Q1 = UInt32(1)
const FLAG_DRAWS = UInt32(2)
const FLAG_SUMMARY = UInt32(4)
const FLAG_SMOOTHED_MEAN = UInt32(8)
const FLAG_LOGLIK = UInt32(16)

# struct hmcgpu_problem (include/hmcgpu.h)
struct Problem
    y::Ptr{Float64}; y_len::Int64; n_series::Int32; n_windows::Int32
    win_series::Ptr{Int32}; win_start::Ptr{Int32}; win_end::Ptr{Int32}; win_id::Ptr{Int64}
    K::Int32; n_chains::Int32; burnin::Int64; nrun::Int64; seed::UInt64
    xi::Ptr{Float64}; alpha::Ptr{Float64}; nu::Ptr{Float64}; beta0::Ptr{Float64}; beta::Ptr{Float64}
    kappa::Float64; is_signal::Ptr{UInt8}; horizons::Ptr{Int32}; n_h::Int32
    X0::Ptr{Int64}; precision::Int32; flags::UInt32
end

# struct hmcgpu_result
mutable struct Result
    mu::Ptr{Float64}; sigma2::Ptr{Float64}; A::Ptr{Float64}; pi_end::Ptr{Float64}; forecasts::Ptr{Float64}; loglik::Ptr{Float64}
    summary_mean::Ptr{Float64}; summary_var::Ptr{Float64}; pib_mean::Ptr{Float64}; insample_forecast_mean::Ptr{Float64}; status::Ptr{Int32}
    gpu_ms::Float64; sweep_kernel_ms::Float64; n_launches::Int64; n_sweep_launches::Int64
    h2d_bytes::Int64; d2h_bytes::Int64; state_steps::Int64
end

struct HmcGpuError <: Exception
    code::Int
    msg::String
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hmcgpu_ctx_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, ref)
        rc == 0 || throw(HmcGpuError(rc, unsafe_string(ccall((:hmcgpu_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        ctx = new(ref[])
        finalizer(c -> (c.h != C_NULL && ccall((:hmcgpu_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        return ctx
    end
end

lasterror(ctx::Context) = unsafe_string(ccall((:hmcgpu_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))

"""
    estimate(ctx, rawdata, win_start, win_end; D, n_chains, burnin, Nrun, seed, horizons, precision)

Many end-date windows in one call (what the SLURM job array of run_hmm.jl does one process at a time).
Returns per-window arrays in the reference's layouts: μ[w] (R×D), σ[w] (R×D), A[w] (R×D×D), πbend[w] (R×D),
forecasts[w] (R×2|H|) with R = n_chains*Nrun (chain-major).  The output buffers are Julia column-major already
(hmcgpu_result), so they are wrapped without a copy.
"""
function estimate(ctx::Context, rawdata::Vector{Float64}, win_start::Vector{Int32}, win_end::Vector{Int32};
                  D::Int = 3, n_chains::Int = 1, burnin::Int = 1_000, Nrun::Int = 1_000, seed::Integer = 1234,
                  horizons::Vector{Int32} = Int32[12], precision::Int = 64)
    nw, nh, R = length(win_start), length(horizons), n_chains * Nrun
    mu = Array{Float64}(undef, R, D, nw); sig = similar(mu); pie = similar(mu)
    A = Array{Float64}(undef, R, D, D, nw)
    fc = Array{Float64}(undef, R, 2nh, nw)
    status = zeros(Int32, n_chains, nw)
    res = Result(pointer(mu), pointer(sig), pointer(A), pointer(pie), nh > 0 ? pointer(fc) : C_NULL, C_NULL,
                 C_NULL, C_NULL, C_NULL, C_NULL, pointer(status), 0.0, 0.0, 0, 0, 0, 0, 0)
    rc = GC.@preserve rawdata win_start win_end horizons mu sig A pie fc status begin
        prob = Problem(pointer(rawdata), length(rawdata), 1, nw, C_NULL, pointer(win_start), pointer(win_end), C_NULL,
                       D, n_chains, burnin, Nrun, UInt64(seed), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 1.0, C_NULL,
                       pointer(horizons), nh, C_NULL, precision, FLAG_REF_Q1 | FLAG_DRAWS)
        ccall((:hmcgpu_estimate, LIB), Cint, (Ptr{Cvoid}, Ref{Problem}, Ref{Result}), ctx.h, prob, res)
    end
    rc < 0 && throw(HmcGpuError(rc, lasterror(ctx)))
    return (μ = mu, σ = sig, A = A, πbend = pie, forecasts = fc, status = status, events = rc, gpu_ms = res.gpu_ms)
end

"""
    estimatemodel(opt; ctx = Context(0), n_chains = 1, precision = 64)

GPU drop-in for `Hmc.estimatemodel(opt)` (src/Hmc.jl:850-865).  `opt` is an `Hmc.estopt`.  `πb` is returned as an
Nrun×1×D array holding the end-of-window row: `saveresults` reads `samples.πb[:, end, :]` (src/Hmc.jl:744), which works
unchanged; the Nrun×N×D tensor of the reference (3.5 GB per end date in production) is never materialised.
"""
function estimatemodel(opt; ctx::Context = Context(0), n_chains::Int = 1, precision::Int = 64)
    isempty(opt.signalRange) || error("signalRange: the noisy-signal tier (estimatesignals!) is not implemented on the GPU path")
    sr = opt.sampleRange
    r = estimate(ctx, Vector{Float64}(opt.rawdata), Int32[first(sr)], Int32[last(sr)]; D = opt.D, n_chains = n_chains,
                 burnin = opt.burnin, Nrun = opt.Nrun, seed = opt.seed, horizons = Int32.(opt.horizons), precision = precision)
    R = n_chains * opt.Nrun
    return (μ = r.μ[:, :, 1], σ = r.σ[:, :, 1], πb = reshape(r.πbend[:, :, 1], R, 1, opt.D), A = r.A[:, :, :, 1],
            forecasts = r.forecasts[:, :, 1], obsdates = fill(opt.dates[opt.endIndex], R))
end

end # module
