# Drop-in for the reference's src/run_model_no_output.jl:3-16.  Same entry point, same return
# value; the step loop runs on the GPU with the state resident in HBM (one upload, one
# `qg_step` for all steps, one download).
include("model.jl")

function run_model_no_output(model::BaroclinicModel)
    zeta, psi = initialise_model(model)
    poisson_chol_fact = get_poisson_cholesky(model.M, model.P, model.dx)                 # plan tokens; the plan
    helmholtz_chol_fact = get_helmholtz_cholesky(model.M, model.P, model.dx, S_eig(model))  # is built in qg_create
    total_steps = floor(Int, model.T / model.dt)
    # f_store = zeros(model.M+2, model.P+2, 2, 3) of the reference lives on the device only

    h = qg_handle(model)
    qg_check(h.ptr, ccall((:qg_upload_initial_state, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                          h.ptr, zeta, psi))
    qg_step!(h, 1, total_steps)          # evolve_zeta! + evolve_psi! for timestep in 1:total_steps
    qg_download!(h, zeta, psi, nothing)

    return zeta, psi
end
