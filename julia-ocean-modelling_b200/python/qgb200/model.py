"""Host-side mirror of the reference's model API (src/model.jl, src/schemes/laplacian.jl,
src/run_model_no_output.jl) over libqgb200.

Same names, same argument order, same error behaviour as the Julia functions, minus the
``!`` that Python identifiers cannot carry.  Arithmetic that the reference performs once per
run on the host (derived parameters, the random initial condition) is done here exactly as
the reference does it; everything that runs once per time step is a CUDA kernel behind the
C ABI.  State arrays are Fortran-ordered ``(M+2, P+2, 2, 3)`` float64 NumPy arrays, byte
compatible with the Julia ``Array{Float64,4}``.

Two ways to drive it:

* the reference's own call pattern — ``evolve_zeta(model, zeta, psi, timestep, f_store)``,
  ``evolve_psi(model, zeta, psi, pchol, hchol)`` — with strict reference semantics (host
  arrays are current after every call; costs one upload + download per call), or
* a :class:`Session`, which keeps the state resident in HBM between steps
  (``run_model_no_output`` uses it: one upload, ``total_steps`` fused steps, one download).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import QGError, qg_params

MINUTES = 60            # src/model.jl:7
DAY = 60 * 60 * 24      # src/model.jl:8
KM = 1000.0             # src/model.jl:9
YEAR = 60 * 60 * 24 * 365  # src/model.jl:10


@dataclass(frozen=True)
class RectangularDomain:
    """src/schemes/laplacian.jl:6-11"""
    x1: float
    x2: float
    y1: float
    y2: float


class BaroclinicModel:
    """src/model.jl:12-34.  Construct with the reference's 15 positional arguments
    ``(H_1, H_2, beta, Lx, Ly, dt, T, U, M, P, dx, visc, r, R_d, initial_kick)``."""

    __slots__ = ("H_1", "H_2", "H", "beta", "Lx", "Ly", "domain", "dt", "T", "U", "M", "P", "dx",
                 "visc", "r", "R_d", "initial_kick")

    def __init__(self, H_1, H_2, beta, Lx, Ly, dt, T, U, M, P, dx, visc, r, R_d, initial_kick):
        s = object.__setattr__
        s(self, "H_1", float(H_1)); s(self, "H_2", float(H_2)); s(self, "H", float(H_1) + float(H_2))
        s(self, "beta", float(beta)); s(self, "Lx", float(Lx)); s(self, "Ly", float(Ly))
        s(self, "domain", RectangularDomain(0.0, float(Lx), 0.0, float(Ly)))
        s(self, "dt", float(dt)); s(self, "T", float(T)); s(self, "U", float(U))
        s(self, "M", int(M)); s(self, "P", int(P)); s(self, "dx", float(dx))
        s(self, "visc", float(visc)); s(self, "r", float(r)); s(self, "R_d", float(R_d))
        s(self, "initial_kick", float(initial_kick))

    def __setattr__(self, k, v):   # immutable like the Julia struct
        raise AttributeError("BaroclinicModel is immutable")

    def _key(self):
        return tuple(getattr(self, k) for k in self.__slots__ if k != "domain")

    def __repr__(self):
        return "BaroclinicModel(" + ", ".join(f"{k}={getattr(self, k)!r}" for k in self.__slots__) + ")"


def ratio_term(model):
    """(f_0/N_0)^2, src/model.jl:109-111"""
    return 0.5 * (model.H_1 + model.H_2) / ((model.R_d * model.R_d) * ((1 / model.H_1) + (1 / model.H_2)))


def S1_plus(model):
    """src/model.jl:113"""
    return (2 * ratio_term(model)) / (model.H_1 * (model.H_1 + model.H_2))


def S2_minus(model):
    """src/model.jl:114"""
    return (2 * ratio_term(model)) / (model.H_2 * (model.H_1 + model.H_2))


def beta_1(model):
    """src/model.jl:117"""
    return model.beta + (S1_plus(model) * model.U)


def beta_2(model):
    """src/model.jl:118"""
    return model.beta - (S2_minus(model) * model.U)


def S_eig(model):
    """src/model.jl:121"""
    return -1 / (model.R_d * model.R_d)


def P_matrix(H_1, H_2):
    """src/model.jl:83-87"""
    P = np.ones((2, 2))
    P[0, 1] = -H_2 / H_1
    return P


def P_inv_matrix(model):
    """src/model.jl:90-99"""
    P = np.zeros((2, 2))
    a = S1_plus(model)
    b = S2_minus(model)
    P[0, 0] = b
    P[0, 1] = a
    P[1, 0] = -b
    P[1, 1] = b
    return (1 / (a + b)) * P


def update_doubly_periodic_bc(b):
    """src/schemes/boundary_conditions.jl:2-13 (host arrays, used for the initial condition)."""
    b[1:-1, 0] = b[1:-1, -2]
    b[1:-1, -1] = b[1:-1, 1]
    b[0, 1:-1] = b[-2, 1:-1]
    b[-1, 1:-1] = b[1, 1:-1]
    b[0, 0] = b[-2, -2]
    b[0, -1] = b[-2, 1]
    b[-1, -1] = b[1, 1]
    b[-1, 0] = b[1, -2]
    return b


def make_params(model):
    """Flatten a model into the C parameter block, evaluating every derived constant with
    the reference's formulas — including P_matrix(H_1, H_1) as evolve_psi! calls it
    (src/model.jl:173)."""
    p = qg_params()
    p.M, p.P = model.M, model.P
    p.dx, p.dt, p.visc, p.r, p.U = model.dx, model.dt, model.visc, model.r, model.U
    p.beta1, p.beta2, p.alpha = beta_1(model), beta_2(model), S_eig(model)
    p.Pinv = (C.c_double * 4)(*P_inv_matrix(model).ravel())
    p.Pfwd = (C.c_double * 4)(*P_matrix(model.H_1, model.H_1).ravel())
    p.H1, p.H2, p.S1 = model.H_1, model.H_2, S1_plus(model)
    return p


def _as_state(a, model, members, name):
    shape = (model.M + 2, model.P + 2, 2, 3) if members == 1 else (model.M + 2, model.P + 2, 2, 3, members)
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.shape != shape or not a.flags.f_contiguous:
        raise ValueError(f"{name} must be a Fortran-ordered float64 array of shape {shape}")
    return a


class Session:
    """Device-resident state of one model (or an ensemble of ``members`` models sharing the
    parameters).  Thin object wrapper over a ``qg_handle*``."""

    def __init__(self, model, members=1, device=0, stream=None):
        self._lib = _lib.load()
        self.model = model
        self.members = int(members)
        self.device = int(device)
        self._h = C.c_void_p()
        params = make_params(model)
        rc = self._lib.qg_create(C.byref(params), self.device, self.members, C.c_void_p(stream or 0),
                                 C.byref(self._h))
        if rc != 0:
            msg = self._lib.qg_last_error(None)
            raise QGError(rc, msg.decode() if msg else "qg_create failed")

    # -- life cycle ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.qg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        _lib.check(self._h, rc)

    @staticmethod
    def _ptr(a):
        return None if a is None else C.c_void_p(a.ctypes.data)

    # -- y-slab decomposition over several GPUs -----------------------------------------
    @staticmethod
    def nccl_unique_id():
        """128-byte NCCL id (create on one rank, broadcast to the others)."""
        buf = C.create_string_buffer(128)
        lib = _lib.load()
        rc = lib.qg_nccl_unique_id(buf)
        if rc != 0:
            raise QGError(rc, lib.qg_last_error(None).decode())
        return buf.raw

    def dist_init(self, rank, nranks, unique_id):
        """Turn this session into rank `rank` of a y-slab decomposition (model.P = local rows)."""
        assert len(unique_id) == 128
        self._ck(self._lib.qg_dist_init(self._h, int(rank), int(nranks), C.c_char_p(unique_id)))

    IPC_BLOB = 256

    def dist_peer_init(self, allgather):
        """Switch the per-step exchanges of a y-slab run from NCCL calls to NVLink peer stores
        (qg_dist_ipc_export / qg_dist_ipc_import).  `allgather(bytes) -> list[bytes]` gathers one
        256-byte blob per rank in rank order (e.g. torch.distributed.all_gather_object).  Returns
        True when the peer path is active, False when the library refused it because two ranks
        share a GPU (the flag barrier needs one rank per GPU): the run then stays on NCCL."""
        buf = C.create_string_buffer(self.IPC_BLOB)
        self._ck(self._lib.qg_dist_ipc_export(self._h, buf))
        blobs = list(allgather(buf.raw))
        assert all(len(b) == self.IPC_BLOB for b in blobs)
        allb = b"".join(blobs)
        if self._lib.qg_dist_ipc_blobs_share_device(C.c_char_p(allb), len(blobs)) == 1:
            return False
        self._ck(self._lib.qg_dist_ipc_import(self._h, C.c_char_p(allb)))
        return True

    # -- state transfer -----------------------------------------------------------------
    def upload(self, zeta=None, psi=None, f_store=None):
        for name, a in (("zeta", zeta), ("psi", psi), ("f_store", f_store)):
            if a is not None:
                _as_state(a, self.model, self.members, name)
        self._ck(self._lib.qg_upload_state(self._h, self._ptr(zeta), self._ptr(psi), self._ptr(f_store)))

    def upload_initial(self, zeta, psi):
        """Level 1 of zeta and psi only; history levels and f_store are zeroed on the device
        (the state initialise_model produces)."""
        for name, a in (("zeta", zeta), ("psi", psi)):
            _as_state(a, self.model, self.members, name)
        self._ck(self._lib.qg_upload_initial_state(self._h, self._ptr(zeta), self._ptr(psi)))

    def upload_initial_raw(self, zeta_ptr, psi_ptr):
        self._ck(self._lib.qg_upload_initial_state(self._h, C.c_void_p(zeta_ptr), C.c_void_p(psi_ptr)))

    def download(self, zeta=None, psi=None, f_store=None):
        for name, a in (("zeta", zeta), ("psi", psi), ("f_store", f_store)):
            if a is not None:
                _as_state(a, self.model, self.members, name)
        self._ck(self._lib.qg_download_state(self._h, self._ptr(zeta), self._ptr(psi), self._ptr(f_store)))

    def init_state(self, seed):
        """initialise_model (src/model.jl:37-62) on the device: seeded Philox noise for psi, q from
        :47-48, zero history.  Nothing crosses PCIe."""
        m = self.model
        assert np.sign(beta_1(m)) == -np.sign(beta_2(m))          # src/model.jl:38
        self._ck(self._lib.qg_init_state(self._h, C.c_uint64(int(seed)), m.initial_kick * m.U * m.Ly,
                                         S1_plus(m), S2_minus(m)))

    def snapshot_begin(self, zeta1=None, psi1=None):
        """Start an asynchronous download of the newest level: ``zeta[:, :, :, 0]`` / ``psi[:, :, :, 0]``
        as F-ordered ``(M+2, P+2, 2[, members])`` arrays (ideally pinned).  Work queued afterwards
        overlaps the copy; the arrays are valid after :meth:`snapshot_end`."""
        shape = (self.model.M + 2, self.model.P + 2, 2) + ((self.members,) if self.members > 1 else ())
        for name, a in (("zeta1", zeta1), ("psi1", psi1)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.shape == shape
                                      and a.flags.f_contiguous and a.flags.writeable):
                raise ValueError(f"{name} must be a writable Fortran-ordered float64 array of shape {shape}")
        self._ck(self._lib.qg_snapshot_begin(self._h, self._ptr(zeta1), self._ptr(psi1)))

    def snapshot_end(self):
        self._ck(self._lib.qg_snapshot_end(self._h))

    def upload_raw(self, zeta_ptr, psi_ptr, f_ptr):
        """Pointers to host buffers in the reference layout (e.g. pinned torch tensors)."""
        self._ck(self._lib.qg_upload_state(self._h, C.c_void_p(zeta_ptr or 0), C.c_void_p(psi_ptr or 0),
                                           C.c_void_p(f_ptr or 0)))

    def download_raw(self, zeta_ptr, psi_ptr, f_ptr):
        self._ck(self._lib.qg_download_state(self._h, C.c_void_p(zeta_ptr or 0), C.c_void_p(psi_ptr or 0),
                                             C.c_void_p(f_ptr or 0)))

    def new_state_array(self):
        shape = (self.model.M + 2, self.model.P + 2, 2, 3)
        if self.members > 1:
            shape += (self.members,)
        return np.zeros(shape, order="F")

    # -- the hot path -------------------------------------------------------------------
    def evolve_zeta(self, timestep):
        self._ck(self._lib.qg_evolve_zeta(self._h, int(timestep)))

    def evolve_psi(self):
        self._ck(self._lib.qg_evolve_psi(self._h))

    def step(self, first_timestep, nsteps):
        self._ck(self._lib.qg_step(self._h, int(first_timestep), int(nsteps)))

    def sync(self):
        self._ck(self._lib.qg_sync(self._h))

    def diagnostics(self):
        e = (C.c_double * self.members)()
        z = (C.c_double * self.members)()
        self._ck(self._lib.qg_diagnostics(self._h, e, z))
        if self.members == 1:
            return e[0], z[0]
        return np.array(e[:]), np.array(z[:])

    EXTREMA = ("max_q1", "min_q1", "max_q2", "min_q2", "max_psi1", "min_psi1", "max_psi2", "min_psi2")

    def extrema(self):
        """Maximum and minimum of q and psi per layer over the interior of the newest level (device
        reduction; what the reference's update_max / update_min, src/run_model.jl:41-53, compute by
        scanning a host matrix).  Returns an array of shape (8,) - or (members, 8) - in the order of
        :attr:`EXTREMA`."""
        out = (C.c_double * (8 * self.members))()
        self._ck(self._lib.qg_extrema(self._h, out))
        a = np.array(out[:]).reshape(self.members, 8)
        return a[0] if self.members == 1 else a

    def solve(self, f, pinned):
        M, P = self.model.M, self.model.P
        f = np.asfortranarray(f, dtype=np.float64)
        if f.shape != (M + 2, P + 2):
            raise ValueError(f"f must have shape {(M + 2, P + 2)}")
        u = np.zeros((M + 2, P + 2), order="F")
        self._ck(self._lib.qg_solve(self._h, 1 if pinned else 0, self._ptr(f), self._ptr(u)))
        return u

    # -- measurement --------------------------------------------------------------------
    def set_profiling(self, enabled):
        self._ck(self._lib.qg_set_profiling(self._h, 1 if enabled else 0))

    def kernel_times(self):
        ms = (C.c_double * _lib.QG_NKERNELS)()
        n = (C.c_int64 * _lib.QG_NKERNELS)()
        self._ck(self._lib.qg_kernel_times(self._h, ms, n))
        return {self._lib.qg_kernel_name(i).decode(): (ms[i], n[i]) for i in range(_lib.QG_NKERNELS)}

    def launch_count(self):
        return int(self._lib.qg_launch_count(self._h))

    def device_layout(self, which):
        base = C.c_void_p()
        v = [C.c_int64() for _ in range(4)]
        self._ck(self._lib.qg_device_layout(self._h, which, C.byref(base), *[C.byref(x) for x in v]))
        return {"base": base.value, "pitch": v[0].value, "xpad": v[1].value, "ypad": v[2].value,
                "field_stride": v[3].value}


# ------------------------------------------------------------------------------------------
# The reference's function API
# ------------------------------------------------------------------------------------------
_sessions = {}


def _session_for(model, device=0):
    key = (model._key(), device)
    s = _sessions.get(key)
    if s is None:
        s = Session(model, 1, device)
        _sessions[key] = s
    return s


def close_sessions():
    for s in _sessions.values():
        s.close()
    _sessions.clear()


class SpectralPlan:
    """What get_poisson_cholesky / get_helmholtz_cholesky return here: an opaque token.  The
    reference returns a CHOLMOD factor (src/schemes/laplacian.jl:60-75); its callers only
    pass it through to evolve_psi!, so the replacement is a description of the operator
    whose plan (FFT twiddles + y-recurrence coefficients) lives inside the device handle."""

    def __init__(self, M, P, dx, alpha, pinned):
        self.M, self.P, self.dx, self.alpha, self.pinned = int(M), int(P), float(dx), float(alpha), bool(pinned)

    def __repr__(self):
        kind = "poisson(pinned)" if self.pinned else f"helmholtz(alpha={self.alpha})"
        return f"SpectralPlan({self.M}x{self.P}, dx={self.dx}, {kind})"


def get_poisson_cholesky(M, P, dx):
    """src/schemes/laplacian.jl:66-75"""
    return SpectralPlan(M, P, dx, 0.0, True)


def get_helmholtz_cholesky(M, P, dx, alpha):
    """src/schemes/laplacian.jl:60-64"""
    return SpectralPlan(M, P, dx, alpha, False)


def initialise_model(model, seed=None, rand_fields=None):
    """src/model.jl:37-62.  ``rand_fields=(r1, r2)`` injects the two uniform [0,1) draws of
    :41-42 (shape (M+2, P+2)); otherwise they come from ``numpy.random.default_rng(seed)``
    (``seed=None``: unseeded, like the reference)."""
    assert np.sign(beta_1(model)) == -np.sign(beta_2(model))
    M, P = model.M, model.P
    if rand_fields is None:
        rng = np.random.default_rng(seed)
        r1 = np.asfortranarray(rng.random((P + 2, M + 2)).T)
        r2 = np.asfortranarray(rng.random((P + 2, M + 2)).T)
    else:
        r1, r2 = (np.asfortranarray(r, dtype=np.float64) for r in rand_fields)
    psi_1 = model.initial_kick * model.U * model.Ly * r1
    psi_2 = model.initial_kick * model.U * model.Ly * r2
    update_doubly_periodic_bc(psi_1)
    update_doubly_periodic_bc(psi_2)
    i = 1.0 / model.dx
    idx2 = i * i

    def lap(u):
        out = np.zeros_like(u, order="F")
        out[1:-1, 1:-1] = (u[:-2, 1:-1] + u[2:, 1:-1] - 4 * u[1:-1, 1:-1] + u[1:-1, :-2] + u[1:-1, 2:]) * idx2
        return update_doubly_periodic_bc(out)

    zeta_1 = lap(psi_1) + S1_plus(model) * (psi_2 - psi_1)
    zeta_2 = lap(psi_2) + S2_minus(model) * (psi_1 - psi_2)
    update_doubly_periodic_bc(zeta_1)
    update_doubly_periodic_bc(zeta_2)
    zeta = np.zeros((M + 2, P + 2, 2, 3), order="F")
    psi = np.zeros((M + 2, P + 2, 2, 3), order="F")
    psi[:, :, 0, 0] = psi_1
    psi[:, :, 1, 0] = psi_2
    zeta[:, :, 0, 0] = zeta_1
    zeta[:, :, 1, 0] = zeta_2
    return zeta, psi


def evolve_zeta(model, zeta, psi, timestep, f_store, device=0):
    """evolve_zeta!(model, zeta, psi, timestep, f_store), src/model.jl:155-158: mutates
    ``zeta`` and ``f_store`` in place."""
    s = _session_for(model, device)
    s.upload(zeta, psi, f_store)
    s.evolve_zeta(timestep)
    s.download(zeta=zeta, f_store=f_store)


def evolve_psi(model, zeta, psi, poisson_cholesky, helmholtz_cholesky, device=0):
    """evolve_psi!(model, zeta, psi, poisson_cholesky, helmholtz_cholesky),
    src/model.jl:172-199: mutates ``psi`` in place."""
    for plan, pinned in ((poisson_cholesky, True), (helmholtz_cholesky, False)):
        if not isinstance(plan, SpectralPlan) or plan.pinned != pinned or (plan.M, plan.P) != (model.M, model.P):
            raise TypeError("evolve_psi expects the plans returned by get_poisson_cholesky / "
                            "get_helmholtz_cholesky for this grid")
    if helmholtz_cholesky.alpha != S_eig(model) or helmholtz_cholesky.dx != model.dx:
        raise ValueError("helmholtz plan was not built with (model.dx, S_eig(model))")
    s = _session_for(model, device)
    s.upload(zeta, psi, None)
    s.evolve_psi()
    s.download(psi=psi)


def run_model_no_output(model, seed=None, rand_fields=None, device=0, total_steps=None, device_ic=None):
    """src/run_model_no_output.jl:3-16: initial condition, plans, ``floor(T/dt)`` steps;
    returns ``(zeta, psi)``.  State stays in HBM for the whole loop.  ``device_ic=<seed>`` draws
    the initial condition on the GPU (:meth:`Session.init_state`) instead of on the host."""
    get_poisson_cholesky(model.M, model.P, model.dx)
    get_helmholtz_cholesky(model.M, model.P, model.dx, S_eig(model))
    if total_steps is None:
        total_steps = int(np.floor(model.T / model.dt))
    with Session(model, 1, device) as s:
        if device_ic is not None:
            s.init_state(device_ic)
            zeta, psi = s.new_state_array(), s.new_state_array()
        else:
            zeta, psi = initialise_model(model, seed=seed, rand_fields=rand_fields)
            s.upload_initial(zeta, psi)  # f_store = zeros (src/run_model_no_output.jl:8) on the device
        s.step(1, total_steps)
        s.download(zeta=zeta, psi=psi)
    return zeta, psi


def sp_solve_poisson(M, P, dx, f, device=0):
    """src/schemes/laplacian.jl:100-111"""
    return _solve(M, P, dx, f, -1.0, True, device)


def sp_solve_modified_helmholtz(M, P, dx, f, alpha, domain=None, device=0):
    """src/schemes/laplacian.jl:78-98: ``f`` is a matrix with ghosts, or a function of (x, y)
    sampled on ``domain`` exactly as the reference does (:89-98)."""
    if callable(f):
        xs = np.linspace(domain.x1 - dx, domain.x2, M + 2)
        ys = np.linspace(domain.y1 - dx, domain.y2, P + 2)
        f = np.array([[f(x, y) for y in ys] for x in xs], dtype=np.float64)
    if not alpha < 0.0:
        raise ValueError("the spectral plan needs alpha < 0 (modified Helmholtz); use sp_solve_poisson for alpha = 0")
    return _solve(M, P, dx, f, alpha, False, device)


def _solve(M, P, dx, f, alpha, pinned, device):
    m = BaroclinicModel(1.0, 1.0, 0.0, M * dx, P * dx, 1.0, 1.0, 0.0, M, P, dx, 0.0, 0.0, 1.0, 0.0)
    lib = _lib.load()
    p = make_params(m)
    p.alpha = alpha
    h = C.c_void_p()
    rc = lib.qg_create(C.byref(p), device, 1, None, C.byref(h))
    if rc != 0:
        raise QGError(rc, lib.qg_last_error(None).decode())
    try:
        f = np.asfortranarray(f, dtype=np.float64)
        u = np.zeros((M + 2, P + 2), order="F")
        _lib.check(h, lib.qg_solve(h, 1 if pinned else 0, C.c_void_p(f.ctypes.data), C.c_void_p(u.ctypes.data)))
    finally:
        lib.qg_destroy(h)
    return u
