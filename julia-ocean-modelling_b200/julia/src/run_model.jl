# Drop-in for the reference's src/run_model.jl:55-93: same entry point `run_model(model, file_name,
# save_results)`, same JLD keys ("zeta_0", "psi_0", "metadata", then "zeta_$timestep" /
# "psi_$timestep" every `sample_timestep` steps, newest level only).  Between samples the state
# stays in HBM; a sample is `qg_snapshot_begin` (packed on the stepping stream, copied to the host
# on a second stream) and the JLD write of sample k overlaps the steps towards sample k+1.
#
# NOTE: no Julia toolchain in the build image; the tested twin is
# julia-ocean-modelling_b200/python/qgb200/runs.py (same call sequence on the same symbols).
using ProgressBars
using JLD

include("model.jl")

function create_metadata(model::BaroclinicModel)      # reference src/run_model.jl:6-20
    sample_interval = 1.0*DAY
    sample_timestep = floor(Int, sample_interval / model.dt)
    total_steps = floor(Int, model.T / model.dt)
    return Dict("dt" => model.dt, "T" => model.T, "sample_interval" => sample_interval,
                "sample_timestep" => sample_timestep, "total_steps" => total_steps)
end

function log_model_params(model::BaroclinicModel)     # reference src/run_model.jl:22-39
    total_steps = floor(Int, model.T / model.dt)
    println("Parameters:")
    println("Lx = ", model.Lx)
    println("Ly = ", model.Ly)
    println("(f_0^2 / N^2): ", ratio_term(model))
    println("S1 = ", S1_plus(model))
    println("S2 = ", S2_minus(model))
    println("Beta_1 = ", beta_1(model))
    println("Beta_2 = ", beta_2(model))
    println("M = ", model.M)
    println("P = ", model.P)
    println("dt = ", model.dt)
    println("T = ", model.T)
    println("U = ", model.U)
    println("Initial kick = ", model.initial_kick)
    println("Total steps = ", total_steps, "\n")
end

# reference src/run_model.jl:41-53 keeps running extrema by scanning host matrices; here the device
# reduces them (qg_extrema: max q1, min q1, max q2, min q2, max psi1, min psi1, max psi2, min psi2)
update_max(current_max::Float64, value::Float64) = value > current_max ? value : current_max
update_min(current_min::Float64, value::Float64) = value < current_min ? value : current_min

function qg_extrema(h)
    out = zeros(8)
    qg_check(h.ptr, ccall((:qg_extrema, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, out))
    return out
end

function qg_diagnostics(h)
    e = zeros(1); z = zeros(1)
    qg_check(h.ptr, ccall((:qg_diagnostics, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, e, z))
    return e[1], z[1]
end

qg_snapshot_begin!(h, zeta1::Array{Float64, 3}, psi1::Array{Float64, 3}) = qg_check(h.ptr,
    ccall((:qg_snapshot_begin, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, zeta1, psi1))
qg_snapshot_end!(h) = qg_check(h.ptr, ccall((:qg_snapshot_end, libqgb200), Cint, (Ptr{Cvoid},), h.ptr))

"""Columns of the `monitor` time series written by `run_model(...; monitor_every = n)`."""
const MONITOR_COLUMNS = ("timestep", "E", "Z", "max_q1", "min_q1", "max_q2", "min_q2",
                         "max_psi1", "min_psi1", "max_psi2", "min_psi2")

# `monitor_every = n` (0 = off, the reference's behaviour) samples energy, enstrophy and the extrema of q and
# psi on the device at step 0 and every n steps, keeps the running extrema with update_max / update_min
# (the reference defines both, src/run_model.jl:41-53, and never calls them) and writes "monitor" (one row
# per sample) and "monitor_running" (8 values) into the output file.
function run_model(model::BaroclinicModel, file_name::String, save_results::Bool; monitor_every::Int=0)
    log_model_params(model)

    sample_interval = 1.0*DAY
    sample_timestep = 2*floor(Int, sample_interval / model.dt)

    poisson_chol_fact = get_poisson_cholesky(model.M, model.P, model.dx)
    helmholtz_chol_fact = get_helmholtz_cholesky(model.M, model.P, model.dx, S_eig(model))
    total_steps = floor(Int, model.T / model.dt)

    zeta, psi = initialise_model(model)
    if save_results
        metadata = create_metadata(model)
        jldopen(file_name, "w") do file
            write(file, "zeta_0", zeta[:,:,:,1])
            write(file, "psi_0", psi[:,:,:,1])
            write(file, "metadata", metadata)
        end
    end

    h = qg_handle(model)
    qg_check(h.ptr, ccall((:qg_upload_initial_state, libqgb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                          h.ptr, zeta, psi))
    snap_zeta = zeros(model.M+2, model.P+2, 2)
    snap_psi = zeros(model.M+2, model.P+2, 2)
    series = Vector{Vector{Float64}}()
    running = Float64[]
    function sample!(t::Int)
        e, z = qg_diagnostics(h)
        ex = qg_extrema(h)
        push!(series, vcat(Float64[t, e, z], ex))
        if isempty(running)
            running = copy(ex)
        else
            for k in 1:8      # odd entries are maxima, even entries minima
                running[k] = isodd(k) ? update_max(running[k], ex[k]) : update_min(running[k], ex[k])
            end
        end
    end
    monitor_every > 0 && sample!(0)

    println("Running simulation... \n")
    t = 0
    pending = 0            # timestep of the snapshot whose copy is still in flight (0 = none)
    bar = ProgressBar(total=total_steps)
    while t < total_steps
        nxt = min(total_steps, (div(t, sample_timestep) + 1) * sample_timestep)
        if monitor_every > 0
            nxt = min(nxt, (div(t, monitor_every) + 1) * monitor_every)
        end
        qg_step!(h, t + 1, nxt - t)              # queued behind the snapshot copy, if any
        if pending > 0                            # write sample k while the GPU steps towards k+1
            qg_snapshot_end!(h)
            jldopen(file_name, "r+") do file
                write(file, "zeta_$pending", snap_zeta)
                write(file, "psi_$pending", snap_psi)
            end
            pending = 0
        end
        update(bar, nxt - t)
        t = nxt
        if monitor_every > 0 && t % monitor_every == 0
            sample!(t)
        end
        if save_results && t % sample_timestep == 0
            qg_snapshot_begin!(h, snap_zeta, snap_psi)
            pending = t
        end
    end
    if pending > 0
        qg_snapshot_end!(h)
        jldopen(file_name, "r+") do file
            write(file, "zeta_$pending", snap_zeta)
            write(file, "psi_$pending", snap_psi)
        end
    end
    qg_download!(h, zeta, psi, nothing)
    if monitor_every > 0 && save_results
        jldopen(file_name, "r+") do file
            write(file, "monitor", permutedims(hcat(series...)))
            write(file, "monitor_running", running)
        end
    end

    return zeta, psi
end
