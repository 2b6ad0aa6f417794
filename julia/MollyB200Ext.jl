# MollyB200Ext.jl — Julia-side shim that plugs libmollyb200.so in behind Molly.jl's own API.
#
# Written against Molly.jl v0.23.3 (src/force.jl, src/energy.jl, src/simulators.jl, src/neighbors.jl,
# ext/MollyCUDAExt.jl). Julia is not installed in the build image, so this file has only been checked by
# reading; the Python mirror `molly.jl_b200/api.py` binds exactly the same C entry points and is what CI runs.
#
# It overrides the same generic functions that ext/MollyCUDAExt.jl overrides (SURVEY.md §8b):
#   Molly.pairwise_forces_loop_gpu!(buffers, sys, pairwise_inters, nbs::Nothing, Val(needs_vir), step_n)   ext:845
#   Molly.pairwise_pe_loop_gpu!(pe_vec_nounits, buffers, sys, pairwise_inters, nbs::Nothing, step_n)        ext:936
#   Molly.simulate!(sys, sim::VelocityVerlet, n_steps; ...)                                                 simulators.jl:547
# (Molly.remove_CM_motion! for CuArray Systems is NOT redefined: the stock extension owns that exact signature)
# and falls through to the stock methods (invoke) for anything it does not recognise: non-cubic boundaries,
# constraints, virtual sites, couplings other than AndersenThermostat, interactions outside
# {LennardJones, Coulomb, CoulombReactionField, CoulombEwald} or unsupported cutoffs / mixing rules.

module MollyB200Ext

using Molly
using CUDA
using Random

const LIB = get(ENV, "MOLLYB200_LIB", joinpath(@__DIR__, "..", "molly.jl_b200", "libmollyb200.so"))

# ---- C structs (include/mollyb200.h) ------------------------------------------------------------------
struct MBInter
    kind::Int32
    cutoff_kind::Int32
    r_cut::Float64
    r_act::Float64
    weight_special::Float64
    coulomb_const::Float64
    solvent_dielectric::Float64
    ewald_alpha::Float64
    sigma_mix::Int32
    eps_mix::Int32
    approx_erfc::Int32
    use_neighbors::Int32
end

struct MBVVParams
    dt::Float64
    n_steps::Int64
    init_step::Int64
    remove_cm_every::Int32
    andersen_kT::Float64
    andersen_prob::Float64
    rng_ctr1::UInt64
    rng_key::UInt64
end

const MB_LJ, MB_COULOMB, MB_CRF, MB_EWALD_REAL = Int32(0), Int32(1), Int32(2), Int32(3)
const MB_CUT_NONE, MB_CUT_DISTANCE, MB_CUT_SHIFTED_POTENTIAL, MB_CUT_SHIFTED_FORCE = Int32(0), Int32(1), Int32(2), Int32(3)
const MB_CUT_CUBIC_SPLINE, MB_CUT_POLYNOMIAL = Int32(4), Int32(5)
const MB_MIX_LORENTZ, MB_MIX_GEOMETRIC = Int32(0), Int32(1)

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:mb_last_error, LIB), Cstring, ()))
    error("libmollyb200 error $rc: $msg")   # same failure mode as ext/MollyCUDAExt.jl:733-739
end

# ---- translation of Molly interaction structs to descriptors -----------------------------------------
# (kind, dist_cutoff, dist_activation)
cutoff_desc(::NoCutoff) = (MB_CUT_NONE, 0.0, 0.0)
cutoff_desc(c::DistanceCutoff) = (MB_CUT_DISTANCE, Float64(ustrip(c.dist_cutoff)), 0.0)
cutoff_desc(c::ShiftedPotentialCutoff) = (MB_CUT_SHIFTED_POTENTIAL, Float64(ustrip(c.dist_cutoff)), 0.0)
cutoff_desc(c::ShiftedForceCutoff) = (MB_CUT_SHIFTED_FORCE, Float64(ustrip(c.dist_cutoff)), 0.0)
cutoff_desc(c::CubicSplineCutoff) = (MB_CUT_CUBIC_SPLINE, Float64(ustrip(c.dist_cutoff)), Float64(ustrip(c.dist_activation)))
cutoff_desc(c::PolynomialCutoff) = (MB_CUT_POLYNOMIAL, Float64(ustrip(c.dist_cutoff)), Float64(ustrip(c.dist_activation)))
cutoff_desc(::Any) = nothing

mix_desc(::Molly.LorentzMixing) = MB_MIX_LORENTZ
mix_desc(::Molly.GeometricMixing) = MB_MIX_GEOMETRIC
mix_desc(::Any) = nothing

function descriptor(inter::LennardJones)
    cd = cutoff_desc(inter.cutoff)
    sm, em = mix_desc(inter.σ_mixing), mix_desc(inter.ϵ_mixing)
    (isnothing(cd) || isnothing(sm) || em != MB_MIX_GEOMETRIC) && return nothing
    !(inter.shortcut isa Molly.LJZeroShortcut) && return nothing
    return MBInter(MB_LJ, cd[1], cd[2], cd[3], Float64(inter.weight_special), 138.93545764, 1.0, 0.0, sm, em, 0,
                   Int32(inter.use_neighbors))
end
function descriptor(inter::Coulomb)
    cd = cutoff_desc(inter.cutoff)
    isnothing(cd) && return nothing
    return MBInter(MB_COULOMB, cd[1], cd[2], cd[3], Float64(inter.weight_special), Float64(ustrip(inter.coulomb_const)),
                   1.0, 0.0, 0, 1, 0, Int32(inter.use_neighbors))
end
function descriptor(inter::CoulombReactionField)
    return MBInter(MB_CRF, MB_CUT_DISTANCE, Float64(ustrip(inter.dist_cutoff)), 0.0, Float64(inter.weight_special),
                   Float64(ustrip(inter.coulomb_const)), Float64(inter.solvent_dielectric), 0.0, 0, 1, 0,
                   Int32(inter.use_neighbors))
end
# CoulombEwald (src/interactions/coulomb.jl:1320-1441): alpha and approximate_erfc are fields of the struct
function descriptor(inter::CoulombEwald)
    return MBInter(MB_EWALD_REAL, MB_CUT_DISTANCE, Float64(ustrip(inter.dist_cutoff)), 0.0, Float64(inter.weight_special),
                   Float64(ustrip(inter.coulomb_const)), 1.0, Float64(ustrip(inter.α)), 0, 1,
                   Int32(inter.approximate_erfc), Int32(inter.use_neighbors))
end
descriptor(::Any) = nothing

# ---- per-System context (what BuffersGPU + GPUNeighborFinder caches hold in the reference) ------------
mutable struct Context
    handle::Ptr{Cvoid}
    cache_generation::Int
end
const CONTEXTS = IdDict{Any, Context}()

function engine_eligible(sys::System{3, <:CuArray, T}, inters) where T
    T in (Float32, Float64) || return nothing
    (sys.boundary isa CubicBoundary || sys.boundary isa TriclinicBoundary{3, <:Any, <:Any, true}) || return nothing
    length(sys.constraints) == 0 || return nothing
    isempty(sys.virtual_sites) || return nothing
    descs = map(descriptor, inters)
    any(isnothing, descs) && return nothing
    return collect(MBInter, descs)
end

function context_for(sys::System{3, <:CuArray, T}, descs::Vector{MBInter}) where T
    nf = sys.neighbor_finder
    gen = nf isa GPUNeighborFinder ? nf.cache_generation : 0
    ctx = get(CONTEXTS, sys.atoms, nothing)
    if isnothing(ctx) || ctx.cache_generation != gen    # exception-list update invalidates cached masks
        isnothing(ctx) || ccall((:mb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx.handle)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        stream = CUDA.stream().handle
        check(ccall((:mb_ctx_create, LIB), Cint, (Cint, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                    CUDA.deviceid(CUDA.device()), T == Float32 ? 32 : 64, stream, h))
        # atoms: Molly's bits layout Atom{Int32,T,T,T,T,T} is read directly from device memory
        check(ccall((:mb_set_atoms, LIB), Cint, (Ptr{Cvoid}, Int64, CuPtr{Cvoid}), h[], length(sys.atoms),
                    pointer(sys.atoms)))
        if sys.boundary isa TriclinicBoundary      # approx_images = true only (engine_eligible); no-list kernel
            bv = sys.boundary.basis_vectors
            basis = Float64[ustrip(bv[i][j]) for i in 1:3 for j in 1:3]   # row-major: bv1, bv2, bv3
            check(ccall((:mb_set_box_triclinic, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h[], basis))
        else
            side = Float64.(ustrip.(sys.boundary.side_lengths))
            check(ccall((:mb_set_box, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h[], collect(side)))
        end
        if nf isa GPUNeighborFinder
            ei, ej = Array(nf.excluded_i), Array(nf.excluded_j)     # sparse 1-based lists, neighbors.jl:104-115
            si, sj = Array(nf.special_i), Array(nf.special_j)
            check(ccall((:mb_set_exceptions, LIB), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Int32}),
                        h[], length(ei), ei, ej, length(si), si, sj))
            check(ccall((:mb_set_neighbor_policy, LIB), Cint, (Ptr{Cvoid}, Float64, Cint), h[],
                        Float64(ustrip(nf.dist_cutoff)), 0))
        end
        ctx = Context(h[], gen)
        CONTEXTS[sys.atoms] = ctx
    end
    check(ccall((:mb_set_inters, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{MBInter}), ctx.handle, length(descs), descs))
    return ctx
end

# ---- forces / energy seam -----------------------------------------------------------------------------
function Molly.pairwise_forces_loop_gpu!(buffers, sys::System{3, <:CuArray, T}, pairwise_inters::Tuple,
                                         nbs::Nothing, ::Val{needs_vir}, step_n) where {T, needs_vir}
    descs = engine_eligible(sys, pairwise_inters)
    if isnothing(descs)
        # the STOCK method's own signature (ext/MollyCUDAExt.jl:845: System{D, <:CuArray, T}, untyped pairwise_inters);
        # naming this method's System{3, ...} signature here would recurse into itself
        return invoke(Molly.pairwise_forces_loop_gpu!,
                      Tuple{Any, System{D, <:CuArray, T} where D, Any, Nothing, Val{needs_vir}, Any},
                      buffers, sys, pairwise_inters, nbs, Val(needs_vir), step_n)
    end
    ctx = context_for(sys, descs)
    vir = needs_vir ? pointer(buffers.virial_nounits) : CU_NULL
    # contract: ADD into buffers.fs_mat (D x N, original order) and buffers.virial_nounits
    check(ccall((:mb_forces, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64),
                ctx.handle, pointer(sys.coords), pointer(buffers.fs_mat), vir, step_n))
    return buffers
end

function Molly.pairwise_pe_loop_gpu!(pe_vec_nounits, buffers, sys::System{3, <:CuArray, T}, pairwise_inters::Tuple,
                                     nbs::Nothing, step_n) where T
    descs = engine_eligible(sys, pairwise_inters)
    if isnothing(descs)
        return invoke(Molly.pairwise_pe_loop_gpu!,   # stock: ext/MollyCUDAExt.jl:936
                      Tuple{Any, Any, System{D, <:CuArray, T} where D, Any, Nothing, Any},
                      pe_vec_nounits, buffers, sys, pairwise_inters, nbs, step_n)
    end
    ctx = context_for(sys, descs)
    check(ccall((:mb_energy, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64),
                ctx.handle, pointer(sys.coords), pointer(pe_vec_nounits), step_n))
    return pe_vec_nounits
end

# ---- specific (bonded) interaction lists -> mb_set_specific ----------------------------------------------
# InteractionList{2,3,4}Atoms of HarmonicBond / HarmonicAngle / PeriodicTorsion (src/types.jl:89-157). Anything else
# makes simulate! fall through to the stock path. A torsion with several Fourier terms becomes one entry per term
# (zero-k padding terms are dropped), like src/interactions/periodic_torsion.jl:100-142 sums them.
function specific_desc(sil::InteractionList2Atoms)
    inters = Array(sil.inters)
    eltype(inters) <: HarmonicBond || return nothing
    idx = Int32.(vcat(Array(sil.is)', Array(sil.js)'))                       # 2 x n, 1-based, column = one term
    par = Float64.(vcat([ustrip(b.k) for b in inters]', [ustrip(b.r0) for b in inters]'))
    return (0, idx, par)
end
function specific_desc(sil::InteractionList3Atoms)
    inters = Array(sil.inters)
    eltype(inters) <: HarmonicAngle || return nothing
    idx = Int32.(vcat(Array(sil.is)', Array(sil.js)', Array(sil.ks)'))
    par = Float64.(vcat([ustrip(a.k) for a in inters]', [ustrip(a.θ0) for a in inters]'))
    return (1, idx, par)
end
function specific_desc(sil::InteractionList4Atoms)
    inters = Array(sil.inters)
    eltype(inters) <: PeriodicTorsion || return nothing
    is, js, ks, ls = Array(sil.is), Array(sil.js), Array(sil.ks), Array(sil.ls)
    idx, par = Int32[], Float64[]
    for (t, tor) in enumerate(inters), m in eachindex(tor.periodicities)
        k = Float64(ustrip(tor.ks[m]))
        k == 0 && continue
        append!(idx, (is[t], js[t], ks[t], ls[t]))
        append!(par, (Float64(tor.periodicities[m]), Float64(ustrip(tor.phases[m])), k))
    end
    return (2, reshape(idx, 4, :), reshape(par, 3, :))
end
specific_desc(::Any) = nothing

function set_specific!(ctx::Context, sys)
    descs = map(specific_desc, sys.specific_inter_lists)
    any(isnothing, descs) && return false
    for (kind, idx, par) in descs     # at most one list per kind (the engine replaces a kind's list on every call)
        check(ccall((:mb_set_specific, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Int32}, Ptr{Float64}),
                    ctx.handle, kind, size(idx, 2), idx, par))
    end
    return true
end

# ---- multi-GPU: one Julia process per GPU (MPI.jl / Distributed); the reference has nothing here -----------
# rank 0: id = comm_unique_id(); broadcast the 128 bytes; every rank: comm_init!(sys, id, rank, nranks).
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:mb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function comm_init!(sys::System{3, <:CuArray}, id::Vector{UInt8}, rank::Integer, nranks::Integer)
    descs = engine_eligible(sys, sys.pairwise_inters)
    isnothing(descs) && error("mollyb200: System is not engine-eligible")
    ctx = context_for(sys, descs)
    check(ccall((:mb_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), ctx.handle, id, rank, nranks))
    return sys
end

# ---- simulate!(sys, ::VelocityVerlet, n) ----------------------------------------------------------------
# Taken over only when nothing but the pairwise path contributes forces and nothing has to run on the host
# every step; loggers fire between chunks of gcd(logger n_steps) steps (SURVEY.md Appendix A.11).
function takeover_params(sys, sim::VelocityVerlet, n_steps, init_step, rng)
    # LJDispersionCorrection adds no force (lennard_jones.jl:252-275); anything else (PME, implicit solvent) -> stock path
    all(gi -> gi isa Molly.LJDispersionCorrection, sys.general_inters) || return nothing
    all(!isnothing, map(specific_desc, sys.specific_inter_lists)) || return nothing
    kT, prob = 0.0, 0.0
    couplings = sim.coupling isa Tuple ? sim.coupling : (sim.coupling,)
    for c in couplings
        c isa Molly.NoCoupling && continue
        c isa AndersenThermostat || return nothing
        kT = Float64(ustrip(sys.k * c.temperature))
        prob = Float64(ustrip(sim.dt / c.coupling_const))
    end
    return MBVVParams(Float64(ustrip(sim.dt)), n_steps, init_step, Int32(sim.remove_CM_motion), kT, prob,
                      rand(rng, UInt64), rand(rng, UInt64))
end

function Molly.simulate!(sys::System{3, <:CuArray, T}, sim::VelocityVerlet, n_steps::Integer;
                         init_step=0, rng=Random.default_rng(), run_loggers=true, kwargs...) where T
    descs = engine_eligible(sys, sys.pairwise_inters)
    p = isnothing(descs) ? nothing : takeover_params(sys, sim, n_steps, init_step, rng)
    if isnothing(p)
        # stock: simulate!(sys, sim::VelocityVerlet, n_steps_or_time; ...) src/simulators.jl:547
        return invoke(Molly.simulate!, Tuple{Any, VelocityVerlet, Any}, sys, sim, n_steps;
                      init_step=init_step, rng=rng, run_loggers=run_loggers, kwargs...)
    end
    ctx = context_for(sys, descs)
    set_specific!(ctx, sys)
    chunk = run_loggers == false || isempty(sys.loggers) ? n_steps :
            max(1, reduce(gcd, (l.n_steps for l in values(sys.loggers))))
    done = 0
    Molly.apply_loggers!(sys, nothing, nothing, init_step, run_loggers)
    while done < n_steps
        m = min(chunk, n_steps - done)
        pp = MBVVParams(p.dt, m, init_step + done, p.remove_cm_every, p.andersen_kT, p.andersen_prob,
                        rand(rng, UInt64), rand(rng, UInt64))
        check(ccall((:mb_simulate_vv, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ref{MBVVParams}),
                    ctx.handle, pointer(sys.coords), pointer(sys.velocities), Ref(pp)))
        done += m
        Molly.apply_loggers!(sys, nothing, nothing, init_step + done, run_loggers)
    end
    return sys
end

# ---- remove_CM_motion! (ext/MollyCUDAExt.jl:2373) ---------------------------------------------------------
# The stock extension's method has exactly the signature System{3, <:CuArray, T}; defining it again would be a method
# overwrite (an error under precompilation). Inside the taken-over simulate! the engine removes the CM motion itself;
# for stand-alone use this module offers its own function instead of replacing Molly's.
function remove_cm_motion!(sys::System{3, <:CuArray, T}) where T
    descs = engine_eligible(sys, sys.pairwise_inters)
    isnothing(descs) && return Molly.remove_CM_motion!(sys)
    ctx = context_for(sys, descs)
    check(ccall((:mb_remove_cm_motion, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}), ctx.handle, pointer(sys.velocities)))
    return sys
end

# The launch-config API of the stock extension (optimize_cuda_launch_config!, src/cuda_config.jl:53, ext:594) is left
# alone: it tunes the stock kernels, which stay the fall-through path; the brick shape of this engine is chosen by the
# library (mb_set_launch_config).

end # module
