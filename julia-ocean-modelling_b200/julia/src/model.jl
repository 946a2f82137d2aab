# Drop-in replacement for the reference's src/model.jl: same names, same signatures, no
# arithmetic beyond the parameter helpers (reference src/model.jl:83-121) and the initial
# condition (reference src/model.jl:37-62).  Everything that runs once per time step is
# forwarded with `ccall` to libqgb200.so (include/qgb200.h), whose CUDA kernels replace
# src/schemes/arakawa.jl, src/schemes/laplacian.jl and src/schemes/boundary_conditions.jl.
#
# NOTE (experimental): there is no Julia toolchain in the build image, so this file has never been
# executed; the tested binding is its Python twin (julia-ocean-modelling_b200/python/qgb200/model.py),
# which binds the same symbols with the same argument order.  tests/test_julia_shim.py parses every
# `ccall` below and checks name, arity and argument types against include/qgb200.h, the field order
# of QGParams against struct qg_params, and that every type is defined before the `ccall` that names it.

using LinearAlgebra
using Libdl

const MINUTES = 60
const DAY = 60*60*24
const KM = 1000.0
const YEAR = 60*60*24*365

const libqgb200 = get(ENV, "QGB200_LIB", joinpath(@__DIR__, "..", "..", "lib", "libqgb200.so"))

# ---- C parameter block (struct qg_params in include/qgb200.h; field order must match) ---------
# Defined before anything is included: the `ccall` signatures in schemes/laplacian.jl name it.
struct QGParams
    M::Int32
    P::Int32
    dx::Float64
    dt::Float64
    visc::Float64
    r::Float64
    U::Float64
    beta1::Float64
    beta2::Float64
    alpha::Float64
    Pinv::NTuple{4,Float64}   # row-major
    Pfwd::NTuple{4,Float64}   # row-major
    H1::Float64
    H2::Float64
    S1::Float64
end

qg_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:qg_last_error, libqgb200), Cstring, (Ptr{Cvoid},), h))

function qg_check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 || error("libqgb200 error $rc: $(qg_error(h))")
    nothing
end

include("schemes/laplacian.jl")            # RectangularDomain, get_*_cholesky, sp_solve_* (shims)
include("host_ic.jl")                       # host-side ghost refresh / Laplacian (initial condition only)

struct BaroclinicModel
    H_1::Float64
    H_2::Float64
    H::Float64
    beta::Float64
    Lx::Float64
    Ly::Float64
    domain::RectangularDomain
    dt::Float64
    T::Float64
    U::Float64
    M::Int
    P::Int
    dx::Float64
    visc::Float64
    r::Float64
    R_d::Float64
    initial_kick::Float64
end

"""Outer constructor, same 15 positional arguments as the reference (src/model.jl:33-34)."""
BaroclinicModel(H_1, H_2, beta, Lx, Ly, dt, T, U, M, P, dx, visc, r, R_d, initial_kick) = BaroclinicModel(
    H_1, H_2, H_1+H_2, beta, Lx, Ly, RectangularDomain(0, Lx, 0, Ly), dt, T, U, M, P, dx, visc, r, R_d, initial_kick)

# ---- derived parameters: verbatim formulas of the reference (src/model.jl:83-121) -----------
P_matrix(H_1::Float64, H_2::Float64) = [1.0 (-H_2 / H_1); 1.0 1.0]

function P_inv_matrix(model::BaroclinicModel)
    a, b = S1_plus(model), S2_minus(model)
    return (1 / (a+b)) * [b a; -b b]      # scalar times matrix, as the reference evaluates it
end

ratio_term(model::BaroclinicModel) = 0.5*(model.H_1 + model.H_2) / ((model.R_d^2) * ((1/model.H_1) + (1/model.H_2)))
S1_plus(model::BaroclinicModel) = (2 * ratio_term(model)) / (model.H_1 * (model.H_1 + model.H_2))
S2_minus(model::BaroclinicModel) = (2 * ratio_term(model)) / (model.H_2 * (model.H_1 + model.H_2))
beta_1(model::BaroclinicModel) = model.beta + (S1_plus(model) * model.U)
beta_2(model::BaroclinicModel) = model.beta - (S2_minus(model) * model.U)
S_eig(model::BaroclinicModel) = -1 / model.R_d^2

# ---- the parameter block of a model: derived constants by the formulas above ----------------
function QGParams(model::BaroclinicModel)
    Pinv = P_inv_matrix(model)
    # evolve_psi! of the reference builds P with (H_1, H_1) (src/model.jl:173); reproduced as it runs
    Pf = P_matrix(model.H_1, model.H_1)
    QGParams(model.M, model.P, model.dx, model.dt, model.visc, model.r, model.U,
             beta_1(model), beta_2(model), S_eig(model),
             (Pinv[1,1], Pinv[1,2], Pinv[2,1], Pinv[2,2]), (Pf[1,1], Pf[1,2], Pf[2,1], Pf[2,2]),
             model.H_1, model.H_2, S1_plus(model))
end

# ---- handle management ------------------------------------------------------------------------
mutable struct QGHandle
    ptr::Ptr{Cvoid}
    model::BaroclinicModel
end

const _handles = Dict{BaroclinicModel,QGHandle}()

function qg_handle(model::BaroclinicModel; device::Integer=0, members::Integer=1)
    get!(_handles, model) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        p = Ref(QGParams(model))
        rc = ccall((:qg_create, libqgb200), Cint, (Ref{QGParams}, Cint, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                   p, device, members, C_NULL, out)
        rc == 0 || error("libqgb200 error $rc: $(qg_error(Ptr{Cvoid}(C_NULL)))")
        h = QGHandle(out[], model)
        finalizer(x -> ccall((:qg_destroy, libqgb200), Cint, (Ptr{Cvoid},), x.ptr), h)
        h
    end
end

_ptr(a::Array{Float64,4}) = pointer(a)
_ptr(::Nothing) = Ptr{Float64}(C_NULL)

qg_upload!(h::QGHandle, zeta, psi, f_store) = qg_check(h.ptr,
    ccall((:qg_upload_state, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          h.ptr, _ptr(zeta), _ptr(psi), _ptr(f_store)))
qg_download!(h::QGHandle, zeta, psi, f_store) = qg_check(h.ptr,
    ccall((:qg_download_state, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          h.ptr, _ptr(zeta), _ptr(psi), _ptr(f_store)))
qg_step!(h::QGHandle, first_timestep::Integer, nsteps::Integer) = qg_check(h.ptr,
    ccall((:qg_step, libqgb200), Cint, (Ptr{Cvoid}, Cint, Cint), h.ptr, first_timestep, nsteps))

# ---- the reference's API --------------------------------------------------------------------------
"""initialise_model on the device (`qg_init_state`): seeded Philox noise instead of the reference's
unseeded `rand`, same arithmetic (reference src/model.jl:37-62); nothing crosses PCIe."""
function initialise_model_device!(h, model::BaroclinicModel, seed::Integer)
    @assert sign(beta_1(model)) == -sign(beta_2(model))
    qg_check(h.ptr, ccall((:qg_init_state, libqgb200), Cint, (Ptr{Cvoid}, UInt64, Cdouble, Cdouble, Cdouble),
                          h.ptr, UInt64(seed), model.initial_kick * model.U * model.Ly, S1_plus(model), S2_minus(model)))
end

"""Initialise the model with a small random psi and then calculate zeta directly
(reference src/model.jl:37-62; host side, once per run)."""
function initialise_model(model::BaroclinicModel)
    @assert sign(beta_1(model)) == -sign(beta_2(model))
    psi_1 = model.initial_kick * model.U * model.Ly * rand(Float64, (model.M+2, model.P+2))
    psi_2 = model.initial_kick * model.U * model.Ly * rand(Float64, (model.M+2, model.P+2))
    update_doubly_periodic_bc!(psi_1)
    update_doubly_periodic_bc!(psi_2)
    zeta_1 = host_laplace_5p(psi_1, model.dx) + S1_plus(model) * (psi_2 - psi_1)
    zeta_2 = host_laplace_5p(psi_2, model.dx) + S2_minus(model) * (psi_1 - psi_2)
    update_doubly_periodic_bc!(zeta_1)
    update_doubly_periodic_bc!(zeta_2)
    zeta = zeros(model.M+2, model.P+2, 2, 3)
    psi = zeros(model.M+2, model.P+2, 2, 3)
    psi[:,:,1,1] = psi_1
    psi[:,:,2,1] = psi_2
    zeta[:,:,1,1] = zeta_1
    zeta[:,:,2,1] = zeta_2
    return zeta, psi
end

"""evolve_zeta!(model, zeta, psi, timestep, f_store) — reference src/model.jl:155-158.
Strict reference semantics: host arrays are current on return (upload, one kernel, download).
Use `run_model_no_output` / `qg_step!` to keep the state resident on the GPU between steps."""
function evolve_zeta!(model::BaroclinicModel, zeta::Array{Float64, 4}, psi::Array{Float64, 4}, timestep::Int, f_store::Array{Float64, 4})
    h = qg_handle(model)
    qg_upload!(h, zeta, psi, f_store)
    qg_check(h.ptr, ccall((:qg_evolve_zeta, libqgb200), Cint, (Ptr{Cvoid}, Cint), h.ptr, timestep))
    qg_download!(h, zeta, nothing, f_store)
end

"""evolve_psi!(model, zeta, psi, poisson_cholesky, helmholtz_cholesky) — reference src/model.jl:172-199.
The two factor arguments are the opaque `SpectralPlan` tokens returned by `get_poisson_cholesky` /
`get_helmholtz_cholesky` (the reference's CHOLMOD factors are only ever passed through)."""
function evolve_psi!(model::BaroclinicModel, zeta::Array{Float64, 4}, psi::Array{Float64, 4}, poisson_cholesky::SpectralPlan, helmholtz_cholesky::SpectralPlan)
    (poisson_cholesky.pinned && !helmholtz_cholesky.pinned) || throw(ArgumentError("plans swapped"))
    h = qg_handle(model)
    qg_upload!(h, zeta, psi, nothing)
    qg_check(h.ptr, ccall((:qg_evolve_psi, libqgb200), Cint, (Ptr{Cvoid},), h.ptr))
    qg_download!(h, nothing, psi, nothing)
end
