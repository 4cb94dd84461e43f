# ExtensibleMCMCCUDA.jl -- reference-side binding of libextmcmc_cuda.so.
#
# UNEXECUTED IN THIS ENVIRONMENT: Julia is not installed in the build image nor on the GPU
# box.  The file is the `ccall` stub a maintainer of ExtensibleMCMC.jl would add; it is
# written against include/extmcmc.h and kept syntactically simple.  The Python mirror in
# extensiblemcmc.jl_b200/ drives the very same symbols through ctypes and is what the
# tests exercise.
#
# Seam used: MCMCBackend (src/types.jl:107-117); MCMC(updates; backend=CUDAMCMCBackend(...))
# (src/mcmc.jl:39-48) makes init! dispatch init_global_workspace(::CUDAMCMCBackend, ...)
# (src/workspaces.jl:38-47) and run! dispatch create_workspace(::CUDAMCMCBackend, ...)
# (src/workspaces.jl:280-287); __run! gets a method for the CUDA workspace (src/run.jl:64).
module ExtensibleMCMCCUDA

using ExtensibleMCMC
const eMCMC = ExtensibleMCMC
const LIB = "libextmcmc_cuda"

struct Adapt
    kind::Int32; adapt_every_k_steps::Int32
    target_accpt_rate::Float64; scale::Float64; min::Float64; max::Float64; offset::Float64
end
struct Update
    kernel::Int32; n_coords::Int32
    coords::Ptr{Int32}; step::Ptr{Float64}; pos::Ptr{UInt8}
    prior::Int32; n_prior_params::Int32; prior_params::Ptr{Float64}
    adapt::Adapt
end
struct Step
    mcmciter::Int64; prev_mcmciter::Int64; pidx::Int32; prev_pidx::Int32
end
struct Config
    abi_version::Int32; device::Int32; n_chains::Int64; chain_offset::Int64
    n_params::Int32; n_updates::Int32; law::Int32; obs_dim::Int32; seed::UInt64
    shard_mode::Int32; rank::Int32; world_size::Int32; history_window::Int32
    roll_window::Int32; use_graphs::Int32; instrument::Int32; sweep_variant::Int32
    stats_mode::Int32; reserved::NTuple{3,Int32}
end

struct CUDAMCMCBackend <: eMCMC.MCMCBackend
    n_chains::Int; device::Int; seed::UInt64; block_len::Int
end
CUDAMCMCBackend(; n_chains=1, device=0, seed=0, block_len=128) =
    CUDAMCMCBackend(n_chains, device, UInt64(seed), block_len)

check(h, rc) = rc == 0 ? nothing :
    error("libextmcmc_cuda: ", unsafe_string(ccall((:extmcmc_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

mutable struct CUDAGlobalWorkspace{T} <: eMCMC.GlobalWorkspace{T}
    handle::Ptr{Cvoid}
    state::Matrix{T}                       # [p, C]
    state_history::Array{T,4}              # [C, p, NU, M] (chain fastest, as the ABI fills it)
    state_proposal_history::Array{T,4}
    ll_history::Array{Float64,3}           # [C, NU, M]
    llprop_history::Array{Float64,3}
    acceptance_history::Array{UInt8,3}
    block_len::Int
    data
    n_local::Int                           # local workspaces created so far (their pidx)
end

# (EXTMCMC_PRIOR_* kind, parameters) of a reference prior (src/priors.jl:18-88); params(dist) of
# Distributions.jl returns exactly the parameter order the ABI documents.
using Distributions
const PRIOR_KIND = Dict(Normal => 2, Gamma => 3, Uniform => 4, Exponential => 6, InverseGamma => 7,
                        Beta => 8, LogNormal => 9, Cauchy => 10)
prior_abi(::eMCMC.ImproperPrior) = (Int32(0), Float64[])
prior_abi(::eMCMC.ImproperPosPrior) = (Int32(1), Float64[])
function prior_abi(pr::eMCMC.StandardPrior)
    for (T, k) in PRIOR_KIND
        pr.dist isa T && return (Int32(k), Float64[params(pr.dist)...])
    end
    error("StandardPrior($(typeof(pr.dist))) not implemented on the GPU path")
end
function prior_abi(pr::eMCMC.ProductPrior)                  # {K, then per factor: kind, dim, p0, p1}
    out = Float64[length(pr.dists)]
    for (dist, idx) in zip(pr.dists, pr.idx)
        k, pp = prior_abi(dist isa eMCMC.Prior ? dist : eMCMC.StandardPrior(dist))
        k == 5 && error("nested ProductPrior not implemented on the GPU path")
        append!(out, (Float64(k), Float64(length(idx)), get(pp, 1, 0.0), get(pp, 2, 0.0)))
    end
    (Int32(5), out)
end
prior_abi(pr) = error("prior $(typeof(pr)) not implemented on the GPU path")

law_id(P::eMCMC.GsnTargetLaw) = length(P.P.μ) == 1 ? Int32(1) : Int32(2)
law_id(P) = error("target law $(typeof(P)) is not implemented on the GPU path")

function eMCMC.init_global_workspace(b::CUDAMCMCBackend, M, updates::Vector{<:eMCMC.MCMCUpdate},
                                     data, θinit::Vector{T}; kwargs...) where T
    p, NU, C = length(θinit), length(updates), b.n_chains
    cfg = Config(1, b.device, C, 0, p, NU, law_id(data.P), length(first(data.obs)), b.seed,
                 0, 0, 1, 2b.block_len, 100, 1, 0, 0, 0, (0, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:extmcmc_create, LIB), Int32, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error("extmcmc_create: ", unsafe_string(ccall((:extmcmc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    for (u, updt) in enumerate(updates)
        rw = updt.rw
        rw isa eMCMC.UniformRandomWalk || error("transition kernel $(typeof(rw)) not implemented on the GPU path")
        coords = Int32.(collect(updt.coords) .- 1)       # 0-based at the ABI
        eps = Float64.(collect(rw.ϵ)); pos = UInt8.(collect(rw.pos))
        prior, pp = prior_abi(updt.prior)
        a = updt.adpt
        adapt = a isa eMCMC.NoAdaptation ? Adapt(0, 100, 0.234, 1.0, 1e-12, 1e7, 1e2) :
                Adapt(1, a.adapt_every_k_steps, a.target_accpt_rate, a.scale, a.min, a.max, a.offset)
        GC.@preserve coords eps pos pp begin
            upd = Update(1, length(coords), pointer(coords), pointer(eps), pointer(pos), prior,
                         length(pp), isempty(pp) ? C_NULL : pointer(pp), adapt)
            check(h[], ccall((:extmcmc_set_update, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Update}), h[], u - 1, upd))
        end
    end
    obs = Float64.(reduce(vcat, data.obs))               # row-major [N][d]
    GC.@preserve obs check(h[], ccall((:extmcmc_upload_obs, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}), h[], obs, length(data.obs), length(first(data.obs)), C_NULL))
    θ0 = repeat(reshape(θinit, 1, p), C, 1)              # [C, p] column-major == chain fastest
    GC.@preserve θ0 check(h[], ccall((:extmcmc_set_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], θ0))
    ws = CUDAGlobalWorkspace{T}(h[], permutedims(θ0), zeros(T, C, p, NU, M), zeros(T, C, p, NU, M),
                                zeros(C, NU, M), zeros(C, NU, M), zeros(UInt8, C, NU, M), b.block_len, data, 0)
    finalizer(w -> ccall((:extmcmc_destroy, LIB), Int32, (Ptr{Cvoid},), w.handle), ws)
    ws
end

# Local workspaces are views into the global arrays; accessors (ll, ll°, accepted, state, ...)
# of src/workspaces.jl:294-385 index them by (mcmciter, pidx).
struct CUDALocalWorkspace{T} <: eMCMC.LocalWorkspace{T}
    gws::CUDAGlobalWorkspace{T}; pidx::Int; name::String
end
# create_workspaces (src/workspaces.jl:362-371) calls this once per update, in schedule order
function eMCMC.create_workspace(::CUDAMCMCBackend, updt, gws::CUDAGlobalWorkspace{T}, M) where T
    gws.n_local += 1
    CUDALocalWorkspace{T}(gws, gws.n_local, string(eMCMC.remove_curly(typeof(updt))))
end
eMCMC.accepted(ws::CUDALocalWorkspace, i::Int) = ws.gws.acceptance_history[:, ws.pidx, i] .!= 0
eMCMC.set_accepted!(ws::CUDALocalWorkspace, i::Int, v) = (ws.gws.acceptance_history[:, ws.pidx, i] .= v)
eMCMC.ll(ws::CUDALocalWorkspace, i::Int) = ws.gws.ll_history[:, ws.pidx, i]
eMCMC.ll°(ws::CUDALocalWorkspace, i::Int) = ws.gws.llprop_history[:, ws.pidx, i]
eMCMC.name_of_update(ws::CUDALocalWorkspace) = ws.name
eMCMC.state(ws::CUDAGlobalWorkspace) = ws.state
eMCMC.state(ws::CUDAGlobalWorkspace, step) = ws.state_history[:, :, step.pidx, step.mcmciter]
eMCMC.state°(ws::CUDAGlobalWorkspace, step) = ws.state_proposal_history[:, :, step.pidx, step.mcmciter]
eMCMC.num_mcmc_steps(ws::CUDAGlobalWorkspace) = size(ws.state_history, 4)
eMCMC.num_updt(ws::CUDAGlobalWorkspace) = size(ws.state_history, 3)

# estim_mean / estim_cov (src/workspaces.jl:121-136): GenericChainStats of every chain,
# mean [C, p], cov [C, p, p] (the ABI fills chain-fastest; entry (a, b) at a + b p)
function chain_stats(ws::CUDAGlobalWorkspace)
    p, C = size(ws.state); NU = eMCMC.num_updt(ws)
    m = zeros(C, p); cv = zeros(C, p * p); ra = zeros(C, NU); na = zeros(Int64, C, NU); np = zeros(Int64, C, NU)
    GC.@preserve m cv ra na np check(ws.handle, ccall((:extmcmc_get_stats, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}), ws.handle, m, cv, ra, na, np))
    (mean = m, cov = reshape(cv, C, p, p), rolling_ar = ra, n_accept = na, n_prop = np)
end
eMCMC.estim_mean(ws::CUDAGlobalWorkspace) = chain_stats(ws).mean
eMCMC.estim_cov(ws::CUDAGlobalWorkspace) = chain_stats(ws).cov

# __run!: walk the schedule on the host, ship blocks; callbacks define the sync points.
function eMCMC.__run!(gws::CUDAGlobalWorkspace, local_wss, updates, schedule, callbacks)
    block = Step[]; seq = Ref(0)
    flush!() = begin
        isempty(block) && return
        check(gws.handle, ccall((:extmcmc_run_block, LIB), Int32, (Ptr{Cvoid}, Ptr{Step}, Int32),
                                gws.handle, block, length(block)))
        fetch_rows!(gws, seq[], block); seq[] += length(block); empty!(block)
    end
    for step in schedule
        pre = [cb for cb in callbacks if eMCMC.check_if_execute(cb, step, eMCMC.__PRESTEP)]
        isempty(pre) || (flush!(); foreach(cb -> eMCMC.execute!(cb, gws, local_wss, step, eMCMC.__PRESTEP), pre))
        push!(block, Step(step.mcmciter, something(step.prev_mcmciter, 0), step.pidx - 1,
                          step.prev_pidx === nothing ? -1 : step.prev_pidx - 1))
        post = [cb for cb in callbacks if eMCMC.check_if_execute(cb, step, eMCMC.__POSTSTEP)]
        if !isempty(post)
            flush!(); foreach(cb -> eMCMC.execute!(cb, gws, local_wss, step, eMCMC.__POSTSTEP), post)
        elseif length(block) >= gws.block_len
            flush!()
        end
    end
    flush!()
    check(gws.handle, ccall((:extmcmc_sync, LIB), Int32, (Ptr{Cvoid},), gws.handle))
end

function fetch_rows!(gws, seq_lo, block)
    n = length(block); C, p = size(gws.state_history, 1), size(gws.state_history, 2)
    th = zeros(C, p, n); thp = zeros(C, p, n); l = zeros(C, n); lp = zeros(C, n); a = zeros(UInt8, C, n)
    GC.@preserve th thp l lp a check(gws.handle, ccall((:extmcmc_get_history, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
        gws.handle, seq_lo, seq_lo + n, th, thp, l, lp, a))
    for (k, s) in enumerate(block)
        i, j = s.mcmciter, s.pidx + 1
        gws.state_history[:, :, j, i] .= th[:, :, k]; gws.state_proposal_history[:, :, j, i] .= thp[:, :, k]
        gws.ll_history[:, j, i] .= l[:, k]; gws.llprop_history[:, j, i] .= lp[:, k]
        gws.acceptance_history[:, j, i] .= a[:, k]
    end
end

end # module
