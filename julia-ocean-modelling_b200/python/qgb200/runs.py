"""run_model with snapshot output (reference src/run_model.jl) and the restart it lacks.

``run_model(model, file_name, save_results)`` keeps the reference's signature and output
schedule: the initial condition as ``zeta_0`` / ``psi_0`` plus a ``metadata`` record, then
``zeta_<t>`` / ``psi_<t>`` (newest level only, ghosts included) every ``sample_timestep`` steps
(src/run_model.jl:58-59, 70-73, 86-90).  The reference writes JLD (HDF5); neither JLD nor h5py
exists here, so the container is an uncompressed ``.npz`` (a zip of ``.npy`` members, appended
to as the run goes) with the same keys.  Between samples the state never leaves the GPU: a
sample is ``qg_snapshot_begin`` (pack on the stepping stream, copy on a second stream), the
next block of steps is queued behind it, and the file write of sample k overlaps the steps
towards sample k+1.

``save_restart`` / ``resume_model`` store and reload the complete stepper state (three levels
of zeta, psi and the RHS history f_store, and the step counter): a resumed run continues bit
for bit as if it had never stopped.
"""
from __future__ import annotations

import json
import os
import zipfile

import numpy as np

from .model import (DAY, BaroclinicModel, S1_plus, S2_minus, S_eig, Session, beta_1, beta_2,
                    get_helmholtz_cholesky, get_poisson_cholesky, initialise_model, ratio_term)


def create_metadata(model):
    """src/run_model.jl:6-20 (the metadata records one-day sampling; the loop itself samples
    every second day, src/run_model.jl:59 — both are kept as they are)."""
    sample_interval = 1.0 * DAY
    return {
        "dt": model.dt,
        "T": model.T,
        "sample_interval": sample_interval,
        "sample_timestep": int(np.floor(sample_interval / model.dt)),
        "total_steps": int(np.floor(model.T / model.dt)),
    }


def log_model_params(model, out=print):
    """src/run_model.jl:22-39"""
    out("Parameters:")
    for k, v in (("Lx", model.Lx), ("Ly", model.Ly), ("(f_0^2 / N^2)", ratio_term(model)),
                 ("S1", S1_plus(model)), ("S2", S2_minus(model)), ("Beta_1", beta_1(model)),
                 ("Beta_2", beta_2(model)), ("M", model.M), ("P", model.P), ("dt", model.dt), ("T", model.T),
                 ("U", model.U), ("Initial kick", model.initial_kick),
                 ("Total steps", int(np.floor(model.T / model.dt)))):
        out(f"{k} = {v}")


def update_max(current_max, value):
    """src/run_model.jl:41-46 (there: the maximum of a host matrix; here the device has reduced it)."""
    return value if value > current_max else current_max


def update_min(current_min, value):
    """src/run_model.jl:48-53"""
    return value if value < current_min else current_min


MONITOR_COLUMNS = ("timestep", "E", "Z") + Session.EXTREMA


def _append(file_name, **arrays):
    """Append ``name -> array`` members to the .npz at `file_name` (created on first use)."""
    with zipfile.ZipFile(file_name, "a", compression=zipfile.ZIP_STORED, allowZip64=True) as zf:
        for name, a in arrays.items():
            with zf.open(name + ".npy", "w", force_zip64=True) as f:
                np.lib.format.write_array(f, np.asanyarray(a), allow_pickle=False)


def _pinned_pair(shape):
    """Two host arrays for a snapshot; page-locked through torch when it is importable (a
    pageable buffer still works, the copy then just does not overlap the following steps)."""
    n = int(np.prod(shape))
    try:
        import torch
        keep = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
        return [k.numpy().reshape(shape, order="F") for k in keep], keep
    except Exception:
        return [np.zeros(shape, order="F") for _ in range(2)], None


def run_model(model, file_name, save_results, seed=None, rand_fields=None, device=0, sample_timestep=None,
              total_steps=None, log=None, monitor_every=None, monitor=None):
    """src/run_model.jl:55-93.  Returns ``(zeta, psi)`` like the reference; `seed`,
    `sample_timestep`, `total_steps` override the reference's unseeded RNG / two-day sampling /
    ``floor(T/dt)`` for tests.

    ``monitor_every=n`` adds the diagnostics suite the reference only sketches (its dead
    update_max / update_min, src/run_model.jl:41-53): at step 0 and every n steps the device reduces
    energy, enstrophy and the extrema of q and psi (18 doubles cross PCIe per sample).  The time
    series - columns :data:`MONITOR_COLUMNS` - and the running extrema over the run are stored in
    the output file as ``monitor`` / ``monitor_running`` and, if a dict is passed as `monitor`,
    returned in it under the same names."""
    if log is not None:
        log_model_params(model, log)
    if sample_timestep is None:
        sample_timestep = 2 * int(np.floor(1.0 * DAY / model.dt))     # src/run_model.jl:58-59
    if total_steps is None:
        total_steps = int(np.floor(model.T / model.dt))
    get_poisson_cholesky(model.M, model.P, model.dx)
    get_helmholtz_cholesky(model.M, model.P, model.dx, S_eig(model))
    zeta, psi = initialise_model(model, seed=seed, rand_fields=rand_fields)
    if save_results:
        if os.path.exists(file_name):
            os.remove(file_name)                                      # jldopen(file_name, "w")
        _append(file_name, zeta_0=zeta[:, :, :, 0], psi_0=psi[:, :, :, 0],
                metadata=np.frombuffer(json.dumps(create_metadata(model)).encode(), dtype=np.uint8))
    (snap_z, snap_p), _keep = _pinned_pair((model.M + 2, model.P + 2, 2))
    series, running = [], None

    def sample(s, t):
        nonlocal running
        E, Z = s.diagnostics()
        ex = s.extrema()
        series.append([float(t), E, Z] + list(ex))
        if running is None:
            running = list(ex)
        else:   # even entries are maxima, odd entries minima
            running = [update_max(r, v) if k % 2 == 0 else update_min(r, v) for k, (r, v) in enumerate(zip(running, ex))]

    with Session(model, 1, device) as s:
        s.upload_initial(zeta, psi)
        if monitor_every:
            sample(s, 0)
        t, pending = 0, None
        while t < total_steps:
            nxt = min(total_steps, (t // sample_timestep + 1) * sample_timestep)
            if monitor_every:
                nxt = min(nxt, (t // monitor_every + 1) * monitor_every)
            s.step(t + 1, nxt - t)                  # queued behind any snapshot copy still in flight
            if pending is not None:                 # write sample k while the GPU steps towards k+1
                s.snapshot_end()
                _append(file_name, **{f"zeta_{pending}": snap_z, f"psi_{pending}": snap_p})
                pending = None
            t = nxt
            if monitor_every and t % monitor_every == 0:
                sample(s, t)
            if save_results and t % sample_timestep == 0:
                s.snapshot_begin(snap_z, snap_p)
                pending = t
        if pending is not None:
            s.snapshot_end()
            _append(file_name, **{f"zeta_{pending}": snap_z, f"psi_{pending}": snap_p})
        s.download(zeta=zeta, psi=psi)
    if monitor_every:
        mon = {"monitor": np.array(series), "monitor_running": np.array(running)}
        if save_results:
            _append(file_name, **mon)
        if monitor is not None:
            monitor.update(mon)
    return zeta, psi


def load_run(file_name):
    """Read a run_model output file: ``(metadata, {timestep: (zeta, psi)})``."""
    with np.load(file_name) as z:
        meta = json.loads(bytes(z["metadata"]).decode())
        steps = sorted(int(k[5:]) for k in z.files if k.startswith("zeta_"))
        return meta, {t: (z[f"zeta_{t}"], z[f"psi_{t}"]) for t in steps}


# ---- restart (absent from the reference: its output holds level 1 only, which cannot restart AB3) ----
_FIELDS = ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx", "visc", "r", "R_d", "initial_kick")


def save_restart(file_name, session, timestep):
    """Write everything the stepper needs to continue after `timestep` completed steps."""
    zeta, psi, f = session.new_state_array(), session.new_state_array(), session.new_state_array()
    session.download(zeta, psi, f)
    m = session.model
    np.savez(file_name, zeta=zeta, psi=psi, f_store=f, timestep=np.int64(timestep),
             model=np.array([float(getattr(m, k)) for k in _FIELDS]), members=np.int64(session.members))


def load_restart(file_name):
    """``(model, zeta, psi, f_store, timestep, members)`` of a save_restart file."""
    with np.load(file_name) as z:
        v = z["model"]
        args = [int(round(x)) if k in ("M", "P") else float(x) for k, x in zip(_FIELDS, v)]
        return (BaroclinicModel(*args), np.asfortranarray(z["zeta"]), np.asfortranarray(z["psi"]),
                np.asfortranarray(z["f_store"]), int(z["timestep"]), int(z["members"]))


def resume_model(file_name, nsteps, device=0):
    """Continue a saved run for `nsteps` more steps; returns ``(zeta, psi, f_store, timestep)``."""
    model, zeta, psi, f, t, members = load_restart(file_name)
    with Session(model, members, device) as s:
        s.upload(zeta, psi, f)
        s.step(t + 1, nsteps)
        s.download(zeta, psi, f)
    return zeta, psi, f, t + nsteps
