# NonlinearSolidsB200.jl -- thin `ccall` shim that plugs libnls_b200.so (include/nls_b200.h) into
# NonlinearSolids.jl's own dispatch points, so an existing script keeps its meshes, materials,
# boundaries, timesteppers and postprocessing and only swaps the solver type:
#
#     solver = B200NewtonSolver(fem; rtol=1e-10, atol=1e-14, maxits=25)     # was NewtonSolver(fem; ...)
#     set_residual_fn!(solver, Residual(mat)); set_jacobian_fn!(solver, Jacobian(mat))
#     ts = PseudoTimeStepper(solver; Δt=0.1, maxtime=1.0); solve!(ts)
#
# NOTE: Julia is not installed in the build image, so this file is written against the C header
# but has never been executed; the identical ABI is exercised from Python (ctypes) by tests/.
# Reference seams it binds (paths in the NonlinearSolids.jl tree):
#   AbstractNonlinearSolver verbs used by timesteppers   src/nonlinearsolvers/newton_fem.jl:74-125
#   solve!(::NewtonSolver, U0)                           src/nonlinearsolvers/newton_fem.jl:132-190
#   compute_internalforce / compute_dinternalforce       src/materials/defaults.jl:29-46, 68-82
#   updateboundaries! / TimeDependent update!            src/fem/fem.jl:68-72, src/fem/boundary.jl:86-92
#   save_state! on poststep!                             src/timesteppers/types.jl:74-78
module NonlinearSolidsB200

using NonlinearSolids
using LinearAlgebra
import NonlinearSolids: solve!, solution, residual, iterations, model, numdof, is_converged,
    set_residual_fn!, set_jacobian_fn!, set_ctx!, compute_internalforce, compute_dinternalforce, save_state!

export B200NewtonSolver, B200Model, NeoHookean, newmark_step!, tangent!, tangent_solve, bordered_solve

const libnls = get(ENV, "NLS_B200_LIB", "libnls_b200.so")

struct NLSError <: Exception
    status::Int32
    msg::String
end

function check(h::Ptr{Cvoid}, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:nls_last_error, libnls), Cstring, (Ptr{Cvoid},), h))
    rc == 1 ? throw(ArgumentError(msg)) : throw(NLSError(rc, msg))   # NLS_ERR_ARG -> ArgumentError
end

"""Compressible Neo-Hookean (builder-defined T2 material; params mu, lambda)."""
Base.@kwdef struct NeoHookean <: NonlinearSolids.AbstractMaterial
    μ::Float64
    λ::Float64
    ρ::Float64 = 1.0
end

"""Small-strain J2 plasticity with linear isotropic hardening on Q1 quads (plane strain) / hexahedra: the stateful-material
protocol of LinearElastoPlasticity1D (trial state in the internal-force pass, stored tangent, `save_state!` on `poststep!`)
on 2-D / 3-D elements (SURVEY 8 f-2; params K, G, sigma_y, H). State arrays: `epsp_xx … epsp_xz`, `alpha` (+ `_n`)."""
Base.@kwdef struct J2Plasticity <: NonlinearSolids.AbstractMaterial
    E::Float64
    ν::Float64
    σy::Float64
    H::Float64
    ρ::Float64 = 1.0
end
NonlinearSolids.has_state(::J2Plasticity) = true

matkind(::LinearElasticity1D) = Int32(0)
matkind(::ExponentialMaterial) = Int32(1)
matkind(::LinearElastoPlasticity1D) = Int32(2)
matkind(::NeoHookean) = Int32(3)
matkind(::J2Plasticity) = Int32(4)
matparams(m::LinearElasticity1D) = [m.E]
matparams(m::ExponentialMaterial) = [m.σ_sat, m.b]
matparams(m::LinearElastoPlasticity1D) = [m.E, m.Hκ, m.Hα, m.κ₀, m.α₀, m.Eep]
matparams(m::NeoHookean) = [m.μ, m.λ]
matparams(m::J2Plasticity) = [m.E / (3 * (1 - 2 * m.ν)), m.E / (2 * (1 + m.ν)), m.σy, m.H]

"""Device-resident mirror of a FEMModel (one GPU). Indices are converted 1-based -> 0-based once."""
mutable struct B200Model
    fem::FEMModel
    h::Ptr{Cvoid}
    dim::Int
    material::Any
    function B200Model(fem::FEMModel; device::Integer=0, dim::Integer=NonlinearSolids.dim(fem))
        href = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:nls_create, libnls), Int32, (Int32, Ref{Ptr{Cvoid}}), device, href)
        rc == 0 || throw(NLSError(rc, unsafe_string(ccall((:nls_last_error, libnls), Cstring, (Ptr{Cvoid},), C_NULL))))
        m = new(fem, href[], dim, nothing)
        finalizer(x -> ccall((:nls_destroy, libnls), Int32, (Ptr{Cvoid},), x.h), m)
        els = elements(fem.mesh)
        nen = length(nodes(els[1]))
        conn = Matrix{Int32}(undef, nen, length(els))                 # column-major = [nel][nen] row-major in C
        for (e, el) in enumerate(els)
            conn[:, e] .= Int32.(nodes(el) .- 1)
        end
        xy = Float64.(permutedims(reshape(collect(coords(fem.mesh)), :, dim)))   # [nn][dim] row-major
        check(m.h, ccall((:nls_set_mesh, libnls), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Int32, Ptr{Int32}, Ptr{Float64}),
                         m.h, dim, length(nodes(fem.mesh)), length(els), nen, conn, xy))
        check(m.h, ccall((:nls_set_discretization, libnls), Int32, (Ptr{Cvoid}, Int32, Int32),
                         m.h, order(fem.P), order(fem.Q)))
        check(m.h, ccall((:nls_build_pattern, libnls), Int32, (Ptr{Cvoid},), m.h))
        m
    end
end

function sync_material!(m::B200Model, mat)
    m.material === mat && return
    p = matparams(mat)
    check(m.h, ccall((:nls_set_material, libnls), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), m.h, matkind(mat), p, length(p)))
    check(m.h, ccall((:nls_set_area, libnls), Int32, (Ptr{Cvoid}, Float64), m.h, Float64(get(m.fem.data, :A, 3e-4))))
    m.material = mat
end

"""Push the current boundary lists/values (after updateboundaries!) to the device."""
function sync_boundaries!(m::B200Model)
    dn, dv, nn, nv = Int64[], Float64[], Int64[], Float64[]
    for bc in m.fem.boundaries
        if isdirichlet(bc)
            append!(dn, nodes(bc) .- 1); append!(dv, values(bc))
        else
            append!(nn, nodes(bc) .- 1); append!(nv, values(bc))
        end
    end
    check(m.h, ccall((:nls_set_dirichlet, libnls), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}), m.h, length(dn), dn, dv))
    check(m.h, ccall((:nls_set_neumann, libnls), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}), m.h, length(nn), nn, nv))
end

"""DistributedForce functors (gallery/distributedforce.jl:3-27): the host evaluates fdist_fn(time(ctx)),
the element integrals run on the device (nls_set_body_force)."""
function sync_external_forces!(m::B200Model, forces, ctx)
    total = zeros(m.dim)
    for f in forces
        f isa DistributedForce || error("only DistributedForce external forces run on the device path")
        fd = f.fdist_fn isa Function ? f.fdist_fn(time(ctx)) : f.fdist_fn
        total .+= fd
    end
    check(m.h, ccall((:nls_set_body_force, libnls), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}),
                     m.h, isempty(forces) ? 0 : m.dim, total))
end

# ---- dispatch point (2): material-level global loops on a device model -------------------------
"""compute_internalforce(material, ::B200Model, U, Fint, ctx; mode): F_int via the element kernel.
`mode=:add` adds to Fint like the reference (defaults.jl:29-46). Pure internal force: `nls_internal_force` leaves U
untouched and subtracts no external load, so the reference's own `Residual` functor (which applies the Dirichlet
values and the external forces itself, residual.jl:12-26) can sit on top of it without counting them twice.
NOTE: this shim has never been executed (no `julia` in the build image); it is derived from include/nls_b200.h."""
function compute_internalforce(material, m::B200Model, U::Vector{Float64}, Fint::Vector{Float64}, ctx; mode=:add)
    mode in (:add, :set) || throw(ArgumentError("Invalid mode: $mode"))
    sync_material!(m, material)
    tmp = similar(Fint)
    check(m.h, ccall((:nls_internal_force, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.h, U, tmp))
    mode == :set ? (Fint .= tmp) : (Fint .+= tmp)
end

"""compute_dinternalforce(material, ::B200Model, U, vals, ctx; mode): CSR values of dF_int/dU."""
function compute_dinternalforce(material, m::B200Model, U::Vector{Float64}, vals::Vector{Float64}, ctx; mode=:add)
    mode in (:add, :set) || throw(ArgumentError("Invalid mode: $mode"))
    sync_material!(m, material)
    tmp = similar(vals)
    check(m.h, ccall((:nls_internal_force_jacobian, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.h, U, tmp))
    mode == :set ? (vals .= tmp) : (vals .+= tmp)
end

# ---- dispatch point (3): the solver verbs the timesteppers consume ------------------------------
struct NewtonOpts
    type_modified::Int32; atol::Float64; rtol::Float64; dtol::Float64; maxits::Int32; lin_rtol::Float64; lin_maxits::Int32
end
mutable struct NewtonResult
    reason::Int32; iterations::Int32; lin_its_total::Int32; nhist::Int32
    norm_r::Float64; norm_r0::Float64; ms_assembly::Float64; ms_linear::Float64; ms_total::Float64
    NewtonResult() = new(0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0)
end

const REASONS = (REASON_NOT_YET_CONVERGED, REASON_CONVERGED_RELATIVE, REASON_CONVERGED_ABSOLUTE,
                 REASON_DIVERGED_MAXITS, REASON_DIVERGED_STEP, REASON_DIVERGED_LINEAR_SOLVE)

"""Sibling of NewtonSolver (its `K::Matrix{Float64}` field is dense, newton_fem.jl:41, so it cannot be reused at
scale). Same keyword fields plus the Krylov controls that replace `\\`."""
Base.@kwdef mutable struct B200NewtonSolver <: AbstractNonlinearSolver
    type::Symbol = :standard
    atol::Float64 = 1e-16
    rtol::Float64 = 1e-16
    dtol::Float64 = 1e6
    maxits::Int = 100
    lin_rtol::Float64 = 1e-12
    lin_maxits::Int = 100_000
    ctx::Any = nothing
    fem::FEMModel
    dev::B200Model
    residual_fn!::Function = (a...) -> nothing
    jacobian_fn!::Function = (a...) -> nothing
    monitor_residual_norm::Bool = false
    monitor_converged_reason::Bool = false
    allow_maxits::Bool = false
    U::Vector{Float64}
    δU::Vector{Float64}
    R::Vector{Float64}
    norm_history::Vector{Float64} = Float64[]
    converged_reason::NonlinearSolids.AbstractConvergedReason = REASON_NOT_YET_CONVERGED
    iterations::Int = 0
    lin_iterations::Int = 0
end

function B200NewtonSolver(fem::FEMModel; device=0, kwargs...)
    n = numdof(fem)
    B200NewtonSolver(; fem=fem, dev=B200Model(fem; device=device), U=zeros(n), δU=zeros(n), R=zeros(n), kwargs...)
end

solution(s::B200NewtonSolver) = s.U
residual(s::B200NewtonSolver) = s.R
iterations(s::B200NewtonSolver) = s.iterations
model(s::B200NewtonSolver) = s.fem
numdof(s::B200NewtonSolver) = numdof(s.fem)
set_ctx!(s::B200NewtonSolver, ctx) = s.ctx = ctx
set_residual_fn!(s::B200NewtonSolver, f::Function) = s.residual_fn! = f
set_jacobian_fn!(s::B200NewtonSolver, f::Function) = s.jacobian_fn! = f
is_converged(s::B200NewtonSolver; allow_maxits=s.allow_maxits) = is_converged(s.converged_reason, allow_maxits=allow_maxits)

"""solve!(solver, U0): the whole loop of newton_fem.jl:132-190 runs on the GPU (nls_newton_solve)."""
function solve!(s::B200NewtonSolver, U0::AbstractVector=zeros(numdof(s.fem)))
    s.residual_fn! isa Residual && s.jacobian_fn! isa Jacobian ||
        error("B200NewtonSolver needs the gallery Residual/Jacobian functors (their material selects the device law)")
    sync_material!(s.dev, s.residual_fn!.material)
    sync_external_forces!(s.dev, s.residual_fn!.externalforces, s.ctx)
    sync_boundaries!(s.dev)                     # TimeDependent values were updated by step!(ts, Δt)
    opts = NewtonOpts(s.type == :modified, s.atol, s.rtol, s.dtol, s.maxits, s.lin_rtol, s.lin_maxits)
    res = NewtonResult()
    hist = zeros(s.maxits + 2)
    u0 = Vector{Float64}(U0)
    check(s.dev.h, ccall((:nls_newton_solve, libnls), Int32,
        (Ptr{Cvoid}, Ref{NewtonOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{NewtonResult}),
        s.dev.h, opts, u0, s.U, s.R, hist, res))
    s.converged_reason = REASONS[res.reason + 1]
    s.iterations = res.iterations
    s.lin_iterations = res.lin_its_total
    s.norm_history = hist[1:res.nhist]
    return
end

# stateful materials: commit the trial state of a converged step. poststep! (types.jl:74-78) calls
# save_state!(material(ts), model(ts), u, ts) with the host FEMModel; this method is more specific in
# its last argument than the reference's (ctx::Any), so it is picked for timestepper-driven runs and
# forwards to the device when the timestepper's solver is a B200NewtonSolver.
function save_state!(mat::LinearElastoPlasticity1D, fem::FEMModel, U, ts::AbstractTimeStepper)
    s = NonlinearSolids.solver(ts)
    if s isa B200NewtonSolver
        check(s.dev.h, ccall((:nls_commit_state, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}), s.dev.h, Vector{Float64}(U)))
    else
        invoke(save_state!, Tuple{LinearElastoPlasticity1D,FEMModel,Any,Any}, mat, fem, U, ts)
    end
end

# ---- dynamics (SURVEY 8f-3): gallery/mass.jl, the dynamic Residual/Jacobian methods, NewmarkTimeStepper ----------
"""Build the device mass once per (density, area): assemblemass! (mass.jl:8-17); `A` defaults to 1.0 there."""
function sync_mass!(m::B200Model, mat)
    check(m.h, ccall((:nls_set_mass, libnls), Int32, (Ptr{Cvoid}, Float64, Float64), m.h,
                     Float64(density(mat)), Float64(get(m.fem.data, :A, 1.0))))
end

"""assemblemass!(material, model, vals): CSR values of M over the model's pattern (K must be zero on entry)."""
function assemblemass!(mat, m::B200Model, vals::Vector{Float64})
    sync_mass!(m, mat)
    check(m.h, ccall((:nls_get_mass, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}), m.h, vals))
end

"""applymass!(material, model, U̇̇, R): R += M U̇̇ (mass.jl:26-37)."""
function applymass!(mat, m::B200Model, Udd::Vector{Float64}, R::Vector{Float64})
    sync_mass!(m, mat)
    check(m.h, ccall((:nls_apply_mass, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.h, Udd, R))
end

"""One Newmark step on the device: the Newton solve on the acceleration with the closures of newmark.jl:75-107,
then update_state! (newmark.jl:114-117). Call from a `solve!(ts::NewmarkTimeStepper)` whose solver is a
B200NewtonSolver, in place of `solve!(ts.solver, ts.U̇̇[step(ts)-1, :])` + `update_state!`:

    a, U, U̇ = newmark_step!(ts.solver, ts.material, ts.β, ts.γ, ts.Δt, ts.Ũ, ts.Ũ̇, ts.U̇̇[step(ts)-1, :])
"""
function newmark_step!(s::B200NewtonSolver, mat, β, γ, Δt, Ut::Vector{Float64}, Udt::Vector{Float64}, Udd0::Vector{Float64})
    sync_material!(s.dev, mat)
    sync_mass!(s.dev, mat)
    sync_external_forces!(s.dev, s.residual_fn!.externalforces, s.ctx)
    sync_boundaries!(s.dev)
    opts = NewtonOpts(s.type == :modified, s.atol, s.rtol, s.dtol, s.maxits, s.lin_rtol, s.lin_maxits)
    res = NewtonResult()
    hist = zeros(s.maxits + 2)
    n = numdof(s.fem)
    a, U, Ud = zeros(n), zeros(n), zeros(n)
    check(s.dev.h, ccall((:nls_newmark_solve, libnls), Int32,
        (Ptr{Cvoid}, Ref{NewtonOpts}, Float64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{NewtonResult}),
        s.dev.h, opts, Float64(β), Float64(γ), Float64(Δt), Ut, Udt, Udd0, a, U, Ud, s.R, hist, res))
    s.converged_reason = REASONS[res.reason + 1]
    s.iterations = res.iterations
    s.lin_iterations = res.lin_its_total
    s.norm_history = hist[1:res.nhist]
    s.U .= a                                  # solution(solver) is the acceleration, as in the reference
    return a, U, Ud
end

# ---- arc-length continuation (SURVEY 8f-4): the bordered solve of newtonarclength_fem.jl:103-104 on the device -------
"""Assemble the tangent at `d` on the device (CSR values stay there) -- the `dFint(fem, d, Kd)` of newtonarclength_fem.jl."""
function tangent!(m::B200Model, mat, d::Vector{Float64})
    sync_material!(m, mat); sync_boundaries!(m)
    check(m.h, ccall((:nls_dev_upload_u, libnls), Int32, (Ptr{Cvoid}, Ptr{Float64}), m.h, d))
    check(m.h, ccall((:nls_dev_assemble, libnls), Int32, (Ptr{Cvoid}, Int32, Int32), m.h, 0, 1))
end

"""`dofs(fem, Kd) \\ dofs(fem, rhs)` on the tangent assembled last; Jacobi-MINRES, valid past limit points where the tangent
is indefinite. `rhs`, result: full-length vectors (zero on constrained DoFs, = expand())."""
function tangent_solve(m::B200Model, rhs::Vector{Float64}; rtol=1e-12, maxits=200_000)
    x = zeros(length(rhs)); its = Ref{Int32}(0); st = Ref{Int32}(0); rr = Ref{Float64}(0.0)
    check(m.h, ccall((:nls_linear_solve_indefinite, libnls), Int32,
        (Ptr{Cvoid}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Float64}, Ref{Int32}),
        m.h, rtol, maxits, rhs, x, its, rr, st))
    st[] == 0 || throw(LinearAlgebra.SingularException(Int(st[])))    # the branch newtonarclength_fem.jl:105-113 catches
    return x
end

"""The bordered update `[Kd -Fext; g_d g_l] \\ [r_d; r_l]` (newtonarclength_fem.jl:103-104) by block elimination:
two solves with the same tangent. All vectors full length; returns (dd, dlambda)."""
function bordered_solve(m::B200Model, r_d, F_ext, g_d, g_l, r_l; kwargs...)
    v1 = tangent_solve(m, r_d; kwargs...); v2 = tangent_solve(m, F_ext; kwargs...)
    dl = (r_l - dot(g_d, v1)) / (g_l + dot(g_d, v2))
    return v1 .+ dl .* v2, dl
end

end # module
