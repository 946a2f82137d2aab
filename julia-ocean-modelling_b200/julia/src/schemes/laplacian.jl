# Shim for the reference's src/schemes/laplacian.jl: the sparse Cholesky factors become opaque
# plan tokens (the spectral plan itself lives inside the libqgb200 handle), and the single-use
# solves forward to qg_solve.

struct RectangularDomain
    x1::Float64
    x2::Float64
    y1::Float64
    y2::Float64
end

"""What get_poisson_cholesky / get_helmholtz_cholesky return instead of a CHOLMOD factor."""
struct SpectralPlan
    M::Int
    P::Int
    dx::Float64
    alpha::Float64
    pinned::Bool
end

get_helmholtz_cholesky(M::Int, P::Int, dx::Float64, alpha::Float64) = SpectralPlan(M, P, dx, alpha, false)  # reference :60-64
get_poisson_cholesky(M::Int, P::Int, dx::Float64) = SpectralPlan(M, P, dx, 0.0, true)                       # reference :66-75

function _qg_single_solve(M::Int, P::Int, dx::Float64, f::Matrix{Float64}, alpha::Float64, pinned::Bool)
    p = Ref(QGParams(Int32(M), Int32(P), dx, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, pinned ? -1.0 : alpha,
                     (1.0, 0.0, 0.0, 1.0), (1.0, 0.0, 0.0, 1.0), 1.0, 1.0, 0.0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:qg_create, libqgb200), Cint, (Ref{QGParams}, Cint, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), p, 0, 1, C_NULL, out)
    rc == 0 || error("libqgb200 error $rc: $(qg_error(Ptr{Cvoid}(C_NULL)))")
    u = zeros(M+2, P+2)
    try
        qg_check(out[], ccall((:qg_solve, libqgb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), out[], pinned ? 1 : 0, f, u))
    finally
        ccall((:qg_destroy, libqgb200), Cint, (Ptr{Cvoid},), out[])
    end
    return u
end

"""reference src/schemes/laplacian.jl:78-86"""
sp_solve_modified_helmholtz(M::Int, P::Int, dx::Float64, f::Matrix{Float64}, alpha::Float64) =
    _qg_single_solve(M, P, dx, f, alpha, false)

"""reference src/schemes/laplacian.jl:89-98"""
function sp_solve_modified_helmholtz(M::Int, P::Int, dx::Float64, f_rhs::Function, alpha::Float64, domain::RectangularDomain)
    xs = range(domain.x1 - dx, domain.x2, length=M+2)
    ys = range(domain.y1 - dx, domain.y2, length=P+2)
    b = [f_rhs(x, y) for x in xs, y in ys]
    return sp_solve_modified_helmholtz(M, P, dx, b, alpha)
end

"""reference src/schemes/laplacian.jl:100-111"""
sp_solve_poisson(M::Int, P::Int, dx::Float64, f::Matrix{Float64}) = _qg_single_solve(M, P, dx, f, 0.0, true)
