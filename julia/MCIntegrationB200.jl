# MCIntegrationB200.jl — the reference-side binding of libmcib200.so (include/mcib200.h).
#
# This is the thin `ccall` shim the north star asks for: host code stays Julia (Configuration, train!-free driver
# loop, Result/average/report come from MCIntegration.jl itself); one iteration's blocks run on the B200 through
# the C ABI.  It replaces the body of `integrate`'s iteration loop (src/main.jl:142-207): `_block!` × block,
# MPIreduce*, doReweightMPI!, train!.
#
# STATUS: written against the header, NOT executed — no Julia runtime exists in the build image (SURVEY §0).  The
# Python mirror (mcintegration.jl_b200/) drives the very same exports and is what the tests run.
module MCIntegrationB200

using MCIntegration
using MCIntegration: Configuration, Continuous, Discrete, CompositeVar, FermiK, Result, Dist

const LIB = get(ENV, "LIBMCIB200", "libmcib200")

# ---- mirror of the POD structs (include/mcib200.h) -------------------------------------------------------------
struct MciLeaf
    kind::Int32; nbins::Int32; adapt::Int32; lower_i::Int32
    lower::Float64; upper::Float64; alpha::Float64
end

struct MciProblem
    solver::Int32; integrand_id::Int32; measure_kind::Int32; n_integrand::Int32; n_topvar::Int32; n_leaf::Int32
    topvar_nleaf::Ptr{Int32}; topvar_offset::Ptr{Int32}; leaf::Ptr{MciLeaf}; dof::Ptr{Int32}; obs_len::Ptr{Int32}
    neighbor_ptr::Ptr{Int32}; neighbor_idx::Ptr{Int32}; userdata::Ptr{Float64}
    n_userdata::Int32; n_block::Int32; neval_per_block::UInt64; seed::UInt64
    measurefreq::Int32; walkers_per_block::Int32; thermal_ratio::Float64; device::Int32; hist_policy::Int32
    jit::Int32; reserved_::Int32; user_source::Cstring       # ABI v2: NVRTC specialisation / user integrand source
    chain_burnin::UInt64                                       # extra unmeasured steps per chain (0 = reference)
    mcmc_start_idx::Int32                                      # ABI v3: `idx` of MCMC.montecarlo (mcmc/montecarlo.jl:76), 1-based
    weight_kind::Int32                                         # ABI v3: 0 = Float64, 1 = ComplexF64 values and observables
end

struct MciIterationOut
    block_obs::Ptr{Float64}; block_norm::Ptr{Float64}; hist::Ptr{Float64}; visited::Ptr{Float64}
    propose::Ptr{Float64}; accept::Ptr{Float64}; neval::Ptr{UInt64}
end

const SOLVER = Dict(:vegas => Int32(0), :vegasmc => Int32(1), :mcmc => Int32(2))

struct MciError <: Exception
    code::Int32; msg::String
end
check(h, rc) = rc == 0 || throw(MciError(rc, unsafe_string(ccall((:mci_last_error, LIB), Cstring, (Ptr{Cvoid},), h))))

leaves(v::Union{Continuous,Discrete}) = [v]
leaves(v::CompositeVar) = collect(v.vars)
leaves(v::FermiK{D}) where {D} = fill(v, D)      # one MCI_LEAF_FERMIK record per Cartesian component; :mcmc only
leafdesc(v::FermiK) = MciLeaf(2, 0, false, 0, v.kF, v.δk, v.maxK)
leafdesc(v::Continuous) = MciLeaf(0, length(v.grid) - 1, v.adapt, 0, v.lower, v.lower + v.range, v.alpha)
leafdesc(v::Discrete) = MciLeaf(1, v.upper - v.lower + 1, v.adapt, v.lower, 0.0, 0.0, v.alpha)

"""
    integrate_b200(integrand::Union{Integer,AbstractString}; userdata=Float64[], measure_kind=0, walkers=0, device=0, jit=0,
                   chain_burnin=1024, comm=nothing, idx=1, debug=false, kwargs...)

Same keyword surface as `MCIntegration.integrate` (src/main.jl:71-90): `solver, config, neval, niter, block, gamma,
adapt, debug, reweight_goal, ignore, measurefreq, thermal_ratio` plus the `Configuration` kwargs (`var, dof, type, obs,
reweight, seed, neighbor`).  The integrand is a built-in device function (`enum mci_integrand_id`) or CUDA C++ source
instead of a closure.  `comm` = an `MPI.Comm` (one process per GPU, like the reference's MPI mode, main.jl:113-122): the
128-byte NCCL id travels over it once, afterwards the library's own all-reduce replaces MPIreduce / MPIbcast.  `idx` =
the integrand an :mcmc chain starts in (mcmc/montecarlo.jl:76).  `type=ComplexF64` (configuration.jl:108) selects
complex integrand values; the source then writes (re, im) pairs.
"""
function integrate_b200(integrand; solver=:vegasmc, config=nothing, neval=1e4, niter=10, block=16,
    gamma=1.0, adapt=true, debug=false, reweight_goal=nothing, ignore::Int=adapt ? 1 : 0, measurefreq=1, thermal_ratio=0.1,
    userdata=Float64[], measure_kind=0, walkers=0, device=0, hist_policy=0, jit=0, chain_burnin=1024, comm=nothing, idx::Int=1,
    kwargs...)
    # `integrand` is a built-in id (enum mci_integrand_id) or a String holding CUDA C++ source that defines
    # `template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w)`
    # (MCI_F_USER = 100, compiled by NVRTC inside mci_create) — the device-side equivalent of the closure f(x, c).
    integrand_id = integrand isa AbstractString ? 100 : Int(integrand)
    source = integrand isa AbstractString ? String(integrand) : ""
    haskey(SOLVER, solver) || error("Solver $solver is not supported!")            # main.jl:263
    config = isnothing(config) ? Configuration(; kwargs...) : config                # main.jl:95-97
    nevalperblock, block = MCIntegration._standardize_block(neval, block)          # main.jl:121 (pads block up to nprocs)
    nprocs = isnothing(comm) ? 1 : _mpi().Comm_size(comm)
    @assert block % nprocs == 0                                                      # main.jl:122
    T = eltype(config.observable) <: AbstractArray ? eltype(eltype(config.observable)) : eltype(config.observable)
    cplx = T <: Complex                                                              # Configuration(type=ComplexF64)
    # walkers = 0: the library fills the GPU with whole CTAs of chains of >= 512 steps (mci_chain_info); walkers = 1 and
    # chain_burnin = 0 reproduce the reference's one-chain-per-block decomposition (DESIGN.md 5.3)
    W = solver == :vegas ? 1 : walkers
    W > 0 && (nevalperblock = (nevalperblock ÷ W) * W)

    lv = vcat((leaves(v) for v in config.var)...)
    topn = Int32[length(leaves(v)) for v in config.var]
    topoff = Int32[v.offset for v in config.var]
    desc = MciLeaf[leafdesc(l) for l in lv]
    N, Nv = config.N, length(config.var)
    dof = Int32[config.dof[i][v] for i in 1:N for v in 1:Nv]                        # row-major [Ni][Nv]
    obslen = Int32[length(o) for o in config.observable]                          # in (complex) entries
    goalbuf = isnothing(reweight_goal) ? Float64[] : Float64.(reweight_goal)        # rooted for the whole call
    nbptr = Int32[0]; nbidx = Int32[]
    for nb in config.neighbor
        append!(nbidx, Int32.(nb .- 1)); push!(nbptr, length(nbidx))
    end
    ud = Float64.(userdata)
    seed = UInt64(rand(config.rng, 1:10^6))                                          # main.jl:136

    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve topn topoff desc dof obslen nbptr nbidx ud source begin
        prob = MciProblem(SOLVER[solver], integrand_id, measure_kind, N, Nv, length(lv), pointer(topn), pointer(topoff),
            pointer(desc), pointer(dof), pointer(obslen), pointer(nbptr), pointer(nbidx), pointer(ud), length(ud), block,
            nevalperblock, seed, measurefreq, W, thermal_ratio, device, hist_policy, jit, 0,
            isempty(source) ? Cstring(C_NULL) : Base.unsafe_convert(Cstring, source), UInt64(chain_burnin),
            Int32(idx), Int32(cplx ? 1 : 0))
        rc = ccall((:mci_create, LIB), Cint, (Ref{MciProblem}, Ref{Ptr{Cvoid}}), prob, h)
        rc == 0 || throw(MciError(rc, unsafe_string(ccall((:mci_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    end
    handle = h[]
    try
        isnothing(comm) || comm_init!(handle, comm)   # blocks b mod nprocs == rank; one ncclAllReduce per iteration
        # warm restart: push the grids of `config` (main.jl:95-97, docs/src/index.md:129-150)
        grids = vcat((l.grid for l in lv if l isa Continuous)..., Float64[])
        dist = vcat((l.distribution for l in lv if l isa Discrete)..., Float64[])
        accum = vcat((l.accumulation for l in lv if l isa Discrete)..., Float64[])
        check(handle, ccall((:mci_set_maps, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            handle, grids, dist, accum, config.reweight))

        Nd = N + 1
        obstot = sum(obslen) * (cplx ? 2 : 1)       # doubles per block: a complex entry is (re, im)
        histlen = sum(d.nbins for d in desc)
        pa = max(Nd, Nv)
        block_obs = zeros(obstot, block)          # C row-major [block][obstot] == Julia column-major (obstot, block)
        block_norm = zeros(block); hist = zeros(histlen); visited = zeros(Nd)
        propose = zeros(pa, Nd, 3); accept = zeros(pa, Nd, 3)  # C [3][Nd][pa] == Julia (pa, Nd, 3): permute for config.propose
        nev = Ref{UInt64}(0)
        results = []
        for iter in 1:niter                                                        # main.jl:142
            GC.@preserve block_obs block_norm hist visited propose accept begin
                out = MciIterationOut(pointer(block_obs), pointer(block_norm), pointer(hist), pointer(visited),
                    pointer(propose), pointer(accept), Base.unsafe_convert(Ptr{UInt64}, nev))
                check(handle, ccall((:mci_run_iteration, LIB), Cint, (Ptr{Cvoid}, UInt64, Ref{MciIterationOut}), handle, iter - 1, out))
            end
            config.neval = nev[]; config.visited .= visited
            config.propose .= permutedims(propose, (3, 2, 1)); config.accept .= permutedims(accept, (3, 2, 1))
            config.normalization = sum(block_norm)
            if debug                                                               # vegas/montecarlo.jl:176-178 etc.
                bad = Ref{UInt64}(0)
                ccall((:mci_nonfinite_count, LIB), Cint, (Ptr{Cvoid}, Ref{UInt64}), handle, bad)
                bad[] > 0 && @warn("integrand is not finite at $(bad[]) evaluation(s) of iteration $iter")
            end
            if solver != :vegas                                                    # main.jl:183
                GC.@preserve goalbuf begin
                    goal = isempty(goalbuf) ? Ptr{Float64}(C_NULL) : pointer(goalbuf)
                    check(handle, ccall((:mci_reweight, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), handle, gamma, goal))
                end
            end
            adapt && check(handle, ccall((:mci_refine, LIB), Cint, (Ptr{Cvoid},), handle))   # main.jl:193-198
            # per-block means -> _mean_std (main.jl:275-287, 296-320)
            m = block_obs ./ reshape(block_norm, 1, :)
            if cplx     # (re, im) doubles -> ComplexF64; squares of the two parts kept separately (main.jl:281,285)
                mc = m[1:2:end, :] .+ m[2:2:end, :] .* im
                sq = real.(mc) .^ 2 .+ (imag.(mc) .^ 2) .* im
                flatSum = vec(sum(mc, dims=2)); flatSq = vec(sum(sq, dims=2))
            else
                flatSum = vec(sum(m, dims=2)); flatSq = vec(sum(m .^ 2, dims=2))
            end
            # regroup the flat entries per integrand like config.observable (scalar or vector observables)
            obsSum = Any[]; obsSq = Any[]; o = 0
            for (i, ob) in enumerate(config.observable)
                n = length(ob)
                push!(obsSum, ob isa AbstractArray ? flatSum[o+1:o+n] : flatSum[o+1])
                push!(obsSq, ob isa AbstractArray ? flatSq[o+1:o+n] : flatSq[o+1]); o += n
            end
            mean, std = MCIntegration._mean_std(obsSum, obsSq, block)
            push!(results, (mean, std, config))
        end
        # pull the trained maps back into `config` so that `config=result.config` keeps working
        check(handle, ccall((:mci_get_maps, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            handle, grids, dist, accum, config.reweight))
        go = 0; dd = 0; da = 0
        for l in lv
            if l isa Continuous
                n = length(l.grid); l.grid .= grids[go+1:go+n]; go += n
            elseif l isa Discrete                                                  # a FermiK has no map to pull back
                n = length(l.distribution); l.distribution .= dist[dd+1:dd+n]; l.accumulation .= accum[da+1:da+n+1]
                dd += n; da += n + 1
            end
        end
        return Result(results, ignore)                                             # statistics.jl:24-55
    finally
        ccall((:mci_destroy, LIB), Cvoid, (Ptr{Cvoid},), handle)
    end
end

_mpi() = Base.require(Base.PkgId(Base.UUID("da04e1cc-30fd-572f-bb4f-1f8673147195"), "MPI"))

# one NCCL unique id per MPI communicator and process: the library keeps the communicator of an id alive, so every
# integrate_b200 call after the first joins it instead of paying NCCL's start-up again
const NCCL_IDS = Dict{Any,Vector{UInt8}}()

"Attach an NCCL communicator: rank 0 creates the id, MPI carries the 128 bytes (replaces src/utility/parallel.jl)."
function comm_init!(handle, comm)
    MPI = _mpi()
    id = get!(NCCL_IDS, comm) do
        fresh = zeros(UInt8, 128)
        MPI.Comm_rank(comm) == 0 && ccall((:mci_comm_unique_id, LIB), Cint, (Ptr{UInt8},), fresh)
        MPI.Bcast!(fresh, 0, comm)
        fresh
    end
    check(handle, ccall((:mci_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), handle, id, MPI.Comm_rank(comm), MPI.Comm_size(comm)))
end

end # module
