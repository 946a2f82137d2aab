# Host-side helpers used only while building the initial condition (once per run); on the
# device the kernels maintain the ghost cells themselves.

"""Periodic ghost refresh of a (M+2)x(P+2) field: every ghost cell takes the value of the
interior cell one period away (semantics of the reference's update_doubly_periodic_bc!,
src/schemes/boundary_conditions.jl:2-13, including the four corners)."""
function update_doubly_periodic_bc!(b::Matrix{Float64})
    W, H = size(b)
    wrap(k, n) = k == 1 ? n - 1 : (k == n ? 2 : k)
    for j in 1:H, i in 1:W
        if i == 1 || i == W || j == 1 || j == H
            b[i, j] = b[wrap(i, W), wrap(j, H)]
        end
    end
    return b
end

"""Five-point Laplacian on the interior followed by the ghost refresh
(reference src/schemes/laplacian.jl:15-27; same summation order)."""
function host_laplace_5p(u::Matrix{Float64}, dx::Float64)
    W, H = size(u)
    out = zeros(W, H)
    for j in 2:H-1, i in 2:W-1
        out[i, j] = (u[i-1, j] + u[i+1, j] - 4u[i, j] + u[i, j-1] + u[i, j+1]) * dx^-2
    end
    return update_doubly_periodic_bc!(out)
end
