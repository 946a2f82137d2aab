"""CPU oracle for the Phillips two-layer QG time-stepping hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs may
import it, and only as the checker.  The product path is the CUDA library behind
``include/qgb200.h`` and fails loudly if that library is missing.

What this is: a NumPy/SciPy restatement of the reference's algorithm, function by function
(every function cites the reference file:line it follows; paths relative to the reference
repo root).  The reference is Julia and cannot run in this image (no ``julia`` binary), and
its only native arithmetic is SuiteSparse CHOLMOD (reached through Julia's ``SparseArrays``
stdlib, version unpinned: no Project.toml/Manifest in the reference).  CHOLMOD solves the SPD
systems assembled in ``src/schemes/laplacian.jl:54-75`` exactly up to round-off, so the
oracle's *direct* back-end assembles the identical matrices and solves them with SuperLU
(``scipy.sparse.linalg.splu``); a second, independent *spectral* back-end (2-D FFT) is
validated against the direct one and used where the direct factorisation does not fit.

Parity pinning: the reference's own known-answer tests for this path (``src/test.jl:8-44``,
``:55-69``, ``:71-103``, ``:105-193``, ``:195-217``, ``:229-238``) are restated in
``tests/test_oracle_reference_fixtures.py`` and pass against this file.  The reference holds
no golden trajectory and cannot be run here, so beyond those unit-level known answers this
oracle is **parity unpinned**: multi-step agreement is with this restatement, not with outputs
of the reference (DESIGN.md section 4 says the same).  What stands in for the missing
trajectory: three independent inversions agreeing to 1e-12, and one whole-step known answer
derived from the equations (a barotropic Rossby wave, ``tests/test_oracle_physics.py``).

Array convention: exactly the reference's.  State arrays are Fortran-ordered
``(M+2, P+2, 2, 3)`` float64 (x index first and contiguous, one ghost ring, layer, time
level; level 0 here = Julia level 1 = newest), so ``arr.ravel(order="K")`` is byte-identical
to the Julia ``Array{Float64,4}``.  Indices below are 0-based: Julia ``[i, j]`` is ``[i-1, j-1]``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MINUTES = 60            # src/model.jl:7
DAY = 60 * 60 * 24      # src/model.jl:8
KM = 1000.0             # src/model.jl:9
YEAR = 60 * 60 * 24 * 365  # src/model.jl:10


# --------------------------------------------------------------------------------------
# Model struct and derived parameters
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RectangularDomain:
    """src/schemes/laplacian.jl:6-11"""
    x1: float
    x2: float
    y1: float
    y2: float


@dataclass(frozen=True)
class BaroclinicModel:
    """src/model.jl:12-30 (inner struct; use :func:`make_model` for the 15-argument form)."""
    H_1: float
    H_2: float
    H: float
    beta: float
    Lx: float
    Ly: float
    domain: RectangularDomain
    dt: float
    T: float
    U: float
    M: int
    P: int
    dx: float
    visc: float
    r: float
    R_d: float
    initial_kick: float


def make_model(H_1, H_2, beta, Lx, Ly, dt, T, U, M, P, dx, visc, r, R_d, initial_kick):
    """Outer constructor, src/model.jl:33-34."""
    return BaroclinicModel(float(H_1), float(H_2), float(H_1) + float(H_2), float(beta), float(Lx),
                           float(Ly), RectangularDomain(0.0, float(Lx), 0.0, float(Ly)), float(dt),
                           float(T), float(U), int(M), int(P), float(dx), float(visc), float(r),
                           float(R_d), float(initial_kick))


def ratio_term(m):
    """(f_0/N_0)^2, src/model.jl:109-111."""
    return 0.5 * (m.H_1 + m.H_2) / ((m.R_d * m.R_d) * ((1 / m.H_1) + (1 / m.H_2)))


def S1_plus(m):
    """src/model.jl:113"""
    return (2 * ratio_term(m)) / (m.H_1 * (m.H_1 + m.H_2))


def S2_minus(m):
    """src/model.jl:114"""
    return (2 * ratio_term(m)) / (m.H_2 * (m.H_1 + m.H_2))


def beta_1(m):
    """src/model.jl:117"""
    return m.beta + (S1_plus(m) * m.U)


def beta_2(m):
    """src/model.jl:118"""
    return m.beta - (S2_minus(m) * m.U)


def S_eig(m):
    """src/model.jl:121"""
    return -1 / (m.R_d * m.R_d)


def P_matrix(H_1, H_2):
    """src/model.jl:83-87"""
    P = np.ones((2, 2))
    P[0, 1] = -H_2 / H_1
    return P


def P_inv_matrix(m):
    """src/model.jl:90-99"""
    P = np.zeros((2, 2))
    a = S1_plus(m)
    b = S2_minus(m)
    P[0, 0] = b
    P[0, 1] = a
    P[1, 0] = -b
    P[1, 1] = b
    return (1 / (a + b)) * P


def _inv_pow2(dx):
    """Julia ``dx^-2`` with a literal exponent lowers to ``i = inv(dx); i*i``."""
    i = 1.0 / dx
    return i * i


# --------------------------------------------------------------------------------------
# Boundary conditions, src/schemes/boundary_conditions.jl
# --------------------------------------------------------------------------------------
def update_doubly_periodic_bc(b):
    """In-place ghost refresh, src/schemes/boundary_conditions.jl:2-13."""
    b[1:-1, 0] = b[1:-1, -2]
    b[1:-1, -1] = b[1:-1, 1]
    b[0, 1:-1] = b[-2, 1:-1]
    b[-1, 1:-1] = b[1, 1:-1]
    b[0, 0] = b[-2, -2]
    b[0, -1] = b[-2, 1]
    b[-1, -1] = b[1, 1]
    b[-1, 0] = b[1, -2]
    return b


def add_doubly_periodic_boundaries(u):
    """src/schemes/boundary_conditions.jl:16-22"""
    M, P = u.shape
    ext = np.zeros((M + 2, P + 2), order="F")
    ext[1:-1, 1:-1] = u
    update_doubly_periodic_bc(ext)
    return ext


# --------------------------------------------------------------------------------------
# Stencil operators
# --------------------------------------------------------------------------------------
def laplace_5p(u, dx):
    """src/schemes/laplacian.jl:15-27 (same summation order)."""
    lap = np.zeros(u.shape, order="F")
    lap[1:-1, 1:-1] = (u[:-2, 1:-1] + u[2:, 1:-1] - 4 * u[1:-1, 1:-1]
                       + u[1:-1, :-2] + u[1:-1, 2:]) * _inv_pow2(dx)
    update_doubly_periodic_bc(lap)
    return lap


def cd(u, dx):
    """x centred difference, src/model.jl:68-80.  ``0.5dx^-1`` parses as ``0.5*(dx^-1)``."""
    out = np.zeros(u.shape, order="F")
    out[1:-1, 1:-1] = (0.5 * (1.0 / dx)) * (u[2:, 1:-1] - u[:-2, 1:-1])
    update_doubly_periodic_bc(out)
    return out


def j_pp(zeta, psi):
    """src/schemes/arakawa.jl:7-20"""
    out = np.zeros(zeta.shape, order="F")
    z, p = zeta, psi
    out[1:-1, 1:-1] = ((z[2:, 1:-1] - z[:-2, 1:-1]) * (p[1:-1, 2:] - p[1:-1, :-2])
                       - (z[1:-1, 2:] - z[1:-1, :-2]) * (p[2:, 1:-1] - p[:-2, 1:-1]))
    return out


def j_pt(zeta, psi):
    """src/schemes/arakawa.jl:22-38"""
    out = np.zeros(zeta.shape, order="F")
    z, p = zeta, psi
    out[1:-1, 1:-1] = (z[2:, 1:-1] * (p[2:, 2:] - p[2:, :-2])
                       - z[:-2, 1:-1] * (p[:-2, 2:] - p[:-2, :-2])
                       - z[1:-1, 2:] * (p[2:, 2:] - p[:-2, 2:])
                       + z[1:-1, :-2] * (p[2:, :-2] - p[:-2, :-2]))
    return out


def j_tp(zeta, psi):
    """src/schemes/arakawa.jl:40-56"""
    out = np.zeros(zeta.shape, order="F")
    z, p = zeta, psi
    out[1:-1, 1:-1] = (z[2:, 2:] * (p[1:-1, 2:] - p[2:, 1:-1])
                       - z[:-2, :-2] * (p[:-2, 1:-1] - p[1:-1, :-2])
                       - z[:-2, 2:] * (p[1:-1, 2:] - p[:-2, 1:-1])
                       + z[2:, :-2] * (p[2:, 1:-1] - p[1:-1, :-2]))
    return out


def J(dx, zeta, psi):
    """Arakawa Jacobian J(zeta, psi), src/schemes/arakawa.jl:58-62."""
    j = (j_pp(zeta, psi) + j_pt(zeta, psi) + j_tp(zeta, psi)) / (3 * 4 * (dx * dx))
    update_doubly_periodic_bc(j)
    return j


# --------------------------------------------------------------------------------------
# Right-hand sides and time stepping, src/model.jl:102-170
# --------------------------------------------------------------------------------------
def store_new_state(arr, new_state, z):
    """src/model.jl:102-106 (z is the 0-based layer)."""
    arr[:, :, z, 2] = arr[:, :, z, 1]
    arr[:, :, z, 1] = arr[:, :, z, 0]
    arr[:, :, z, 0] = new_state


def zeta_f1(m, zeta, psi):
    """src/model.jl:139-145"""
    v_term = m.visc * laplace_5p(laplace_5p(psi, m.dx), m.dx)
    J_term = J(m.dx, zeta, psi)
    beta_term = beta_1(m) * cd(psi, m.dx)
    U_term = m.U * cd(zeta, m.dx)
    return v_term - J_term - beta_term - U_term


def zeta_f2(m, zeta, psi):
    """src/model.jl:147-153"""
    v_term = m.visc * laplace_5p(laplace_5p(psi, m.dx), m.dx)
    J_term = J(m.dx, zeta, psi)
    beta_term = beta_2(m) * cd(psi, m.dx)
    r_term = m.r * laplace_5p(psi, m.dx)
    return v_term - J_term - beta_term - r_term


def eulers_method(m, f, zeta, psi, z, f_store):
    """src/model.jl:123-127"""
    f1 = f(m, zeta[:, :, z, 0].copy(order="F"), psi[:, :, z, 0].copy(order="F"))
    store_new_state(f_store, f1, z)
    return zeta[:, :, z, 0] + (m.dt * f1)


def AB3(m, f, zeta, psi, z, f_store):
    """src/model.jl:129-136"""
    f1 = f(m, zeta[:, :, z, 0].copy(order="F"), psi[:, :, z, 0].copy(order="F"))
    store_new_state(f_store, f1, z)
    f2 = f_store[:, :, z, 1]
    f3 = f_store[:, :, z, 2]
    update = m.dt * ((23 / 12) * f1 - (16 / 12) * f2 + (5 / 12) * f3)
    return zeta[:, :, z, 0] + update


def evolve_zeta_layer(m, zeta, psi, timestep, layer, f, f_store):
    """src/model.jl:160-170 (timestep is 1-based as in the reference)."""
    if timestep == 1 or timestep == 2:
        new_zeta = eulers_method(m, f, zeta, psi, layer, f_store)
    else:
        new_zeta = AB3(m, f, zeta, psi, layer, f_store)
    store_new_state(zeta, new_zeta, layer)


def evolve_zeta(m, zeta, psi, timestep, f_store):
    """src/model.jl:155-158"""
    evolve_zeta_layer(m, zeta, psi, timestep, 0, zeta_f1, f_store)
    evolve_zeta_layer(m, zeta, psi, timestep, 1, zeta_f2, f_store)


# --------------------------------------------------------------------------------------
# Inversion back-ends (stand-ins for the CHOLMOD factors of src/schemes/laplacian.jl:60-75)
# --------------------------------------------------------------------------------------
def laplacian_1d(N):
    """src/schemes/laplacian.jl:30"""
    import scipy.sparse as sp
    return sp.diags([np.ones(N - 1), -2 * np.ones(N), np.ones(N - 1)], [-1, 0, 1], format="lil")


def laplacian_1d_periodic(N):
    """src/schemes/laplacian.jl:40-45"""
    lap = laplacian_1d(N)
    lap[0, N - 1] = 1
    lap[N - 1, 0] = 1
    return lap.tocsc()


def laplacian_2d_doubly_periodic(M, P):
    """src/schemes/laplacian.jl:47-51; unknown index = i + M*j (x fastest)."""
    import scipy.sparse as sp
    Dx = laplacian_1d_periodic(M)
    Dy = laplacian_1d_periodic(P)
    return (sp.kron(sp.identity(P), Dx) + sp.kron(Dy, sp.identity(M))).tocsc()


def construct_spA(M, P, dx, alpha):
    """src/schemes/laplacian.jl:54-58"""
    import scipy.sparse as sp
    A = laplacian_2d_doubly_periodic(M, P)
    A = A + alpha * (dx * dx) * sp.identity(M * P)
    return (_inv_pow2(dx) * A).tocsc()


class DirectFactor:
    """Sparse direct factor of the reference's matrix (SuperLU stands in for CHOLMOD)."""

    def __init__(self, A, M, P):
        from scipy.sparse.linalg import splu
        self.M, self.P = M, P
        self.lu = splu(A.tocsc())

    def solve(self, b):
        return self.lu.solve(b)


def get_helmholtz_cholesky(M, P, dx, alpha):
    """src/schemes/laplacian.jl:60-64"""
    A = -construct_spA(M, P, dx, alpha)
    return DirectFactor(A, M, P)


def get_poisson_cholesky(M, P, dx):
    """src/schemes/laplacian.jl:66-75: first unknown pinned (row/col zeroed, A[0,0]=1)."""
    A = (-construct_spA(M, P, dx, 0.0)).tolil()
    A[:, 0] = 0
    A[0, :] = 0
    A[0, 0] = 1
    return DirectFactor(A.tocsc(), M, P)


class SpectralFactor:
    """Second, independent back-end: exact diagonalisation of the same periodic matrices.

    ``solve(b)`` returns the solution of ``A u = b`` for the matrix ``A`` of
    ``get_helmholtz_cholesky`` (``pinned=False``) or ``get_poisson_cholesky`` (``pinned=True``).
    The circulant 1-D Laplacian ``[-2 1 ... 1]`` (src/test.jl:229-238) has eigenvalues
    ``2cos(2 pi k/N) - 2``.  Pinned Poisson: the pinned system is ``-Lap u = b`` at every node
    but node 0 with ``u[0] = 0``; solvability of the periodic problem fixes the residual at
    node 0, i.e. ``b[0]`` is replaced by ``-sum(b[1:])``, the zero mode is dropped and the
    solution is shifted so that ``u[0] = 0``.
    """

    def __init__(self, M, P, dx, alpha, pinned):
        self.M, self.P, self.pinned = M, P, pinned
        lx = 2 * np.cos(2 * np.pi * np.arange(M) / M) - 2
        ly = 2 * np.cos(2 * np.pi * np.arange(P) / P) - 2
        lam = (lx[:, None] + ly[None, :]) * _inv_pow2(dx) + alpha      # eigenvalues of +A_ref
        self.neg_lam = -lam                                           # A = -construct_spA
        if pinned:
            self.neg_lam[0, 0] = 1.0

    def solve(self, b):
        M, P = self.M, self.P
        rhs = np.array(b, dtype=np.float64).reshape((M, P), order="F")
        if self.pinned:
            rhs[0, 0] = 0.0
            rhs[0, 0] = -rhs.sum()
        bh = np.fft.fft2(rhs)
        if self.pinned:
            bh[0, 0] = 0.0
        u = np.real(np.fft.ifft2(bh / self.neg_lam))
        if self.pinned:
            u = u - u[0, 0]
        return u.ravel(order="F")


def get_helmholtz_spectral(M, P, dx, alpha):
    return SpectralFactor(M, P, dx, alpha, pinned=False)


def get_poisson_spectral(M, P, dx):
    return SpectralFactor(M, P, dx, 0.0, pinned=True)


def sp_solve_modified_helmholtz(M, P, dx, f, alpha, factor=None):
    """src/schemes/laplacian.jl:78-86 (matrix right-hand side incl. ghosts)."""
    fac = factor if factor is not None else get_helmholtz_spectral(M, P, dx, alpha)
    b = -f[1:-1, 1:-1].ravel(order="F")
    u = fac.solve(b).reshape((M, P), order="F")
    return add_doubly_periodic_boundaries(u)


def sp_solve_poisson(M, P, dx, f, factor=None):
    """src/schemes/laplacian.jl:100-111"""
    fac = factor if factor is not None else get_poisson_spectral(M, P, dx)
    b = -f[1:-1, 1:-1].ravel(order="F")
    b[0] = 0
    u = fac.solve(b).reshape((M, P), order="F")
    return add_doubly_periodic_boundaries(u)


def evolve_psi(m, zeta, psi, poisson_factor, helmholtz_factor):
    """src/model.jl:172-199, including the (H_1, H_1) back-projection matrix of :173."""
    Pm = P_matrix(m.H_1, m.H_1)
    P_inv = P_inv_matrix(m)
    zt = [P_inv[i, 0] * zeta[:, :, 0, 0] + P_inv[i, 1] * zeta[:, :, 1, 0] for i in range(2)]

    b = -zt[0][1:-1, 1:-1].ravel(order="F")
    b[0] = 0
    u = poisson_factor.solve(b).reshape((m.M, m.P), order="F")
    new_psi_tilde_1 = add_doubly_periodic_boundaries(u)

    b = -zt[1][1:-1, 1:-1].ravel(order="F")
    u = helmholtz_factor.solve(b).reshape((m.M, m.P), order="F")
    new_psi_tilde_2 = add_doubly_periodic_boundaries(u)

    for i in range(2):
        new_psi = Pm[i, 0] * new_psi_tilde_1 + Pm[i, 1] * new_psi_tilde_2
        store_new_state(psi, new_psi, i)


# --------------------------------------------------------------------------------------
# Initial condition and drivers
# --------------------------------------------------------------------------------------
def seeded_random_fields(m, seed):
    """The two ``rand(Float64, (M+2, P+2))`` draws of src/model.jl:41-42, made reproducible.

    The reference's RNG is unseeded, so parity runs inject these arrays on both sides.
    Drawn in memory order of the Fortran-ordered (M+2, P+2) array, psi_1 first.
    """
    rng = np.random.default_rng(seed)
    r1 = np.asfortranarray(rng.random((m.P + 2, m.M + 2)).T)
    r2 = np.asfortranarray(rng.random((m.P + 2, m.M + 2)).T)
    return r1, r2


def initialise_model(m, seed=1, rand_fields=None):
    """src/model.jl:37-62 with the random draws supplied (seeded) instead of unseeded."""
    b1, b2 = beta_1(m), beta_2(m)
    assert np.sign(b1) == -np.sign(b2), "sign(beta_1) must be -sign(beta_2) (src/model.jl:38)"
    r1, r2 = rand_fields if rand_fields is not None else seeded_random_fields(m, seed)
    psi_1 = m.initial_kick * m.U * m.Ly * r1
    psi_2 = m.initial_kick * m.U * m.Ly * r2
    update_doubly_periodic_bc(psi_1)
    update_doubly_periodic_bc(psi_2)
    zeta_1 = laplace_5p(psi_1, m.dx) + S1_plus(m) * (psi_2 - psi_1)
    zeta_2 = laplace_5p(psi_2, m.dx) + S2_minus(m) * (psi_1 - psi_2)
    update_doubly_periodic_bc(zeta_1)
    update_doubly_periodic_bc(zeta_2)
    zeta = np.zeros((m.M + 2, m.P + 2, 2, 3), order="F")
    psi = np.zeros((m.M + 2, m.P + 2, 2, 3), order="F")
    psi[:, :, 0, 0] = psi_1
    psi[:, :, 1, 0] = psi_2
    zeta[:, :, 0, 0] = zeta_1
    zeta[:, :, 1, 0] = zeta_2
    return zeta, psi


def make_factors(m, backend="direct"):
    """The two factorisations of src/run_model_no_output.jl:5-6."""
    if backend == "direct":
        return (get_poisson_cholesky(m.M, m.P, m.dx),
                get_helmholtz_cholesky(m.M, m.P, m.dx, S_eig(m)))
    if backend == "spectral":
        return (get_poisson_spectral(m.M, m.P, m.dx),
                get_helmholtz_spectral(m.M, m.P, m.dx, S_eig(m)))
    raise ValueError(backend)


def run_steps(m, zeta, psi, f_store, factors, first_timestep, nsteps):
    """Loop body of src/run_model_no_output.jl:10-13 for ``nsteps`` steps (in place)."""
    pf, hf = factors
    for timestep in range(first_timestep, first_timestep + nsteps):
        evolve_zeta(m, zeta, psi, timestep, f_store)
        evolve_psi(m, zeta, psi, pf, hf)


def run_model_no_output(m, seed=1, backend="direct", total_steps=None):
    """src/run_model_no_output.jl:3-16"""
    zeta, psi = initialise_model(m, seed)
    factors = make_factors(m, backend)
    if total_steps is None:
        total_steps = int(np.floor(m.T / m.dt))
    f_store = np.zeros((m.M + 2, m.P + 2, 2, 3), order="F")
    run_steps(m, zeta, psi, f_store, factors, 1, total_steps)
    return zeta, psi


# --------------------------------------------------------------------------------------
# Diagnostics (the reference defines none; SURVEY.md App. A.6 fixes the definition that
# both the oracle and the CUDA path use for the 1000-step energy / enstrophy gate)
# --------------------------------------------------------------------------------------
def diagnostics(m, zeta, psi):
    """Domain-integrated energy E and enstrophy Z of time level 0 (float64)."""
    d = m.dx
    a = S1_plus(m)
    E = 0.0
    Z = 0.0
    Hs = (m.H_1, m.H_2)
    for layer in range(2):
        p = psi[:, :, layer, 0]
        q = zeta[:, :, layer, 0]
        ux = (0.5 * (1.0 / d)) * (p[2:, 1:-1] - p[:-2, 1:-1])
        uy = (0.5 * (1.0 / d)) * (p[1:-1, 2:] - p[1:-1, :-2])
        E += 0.5 * d * d * Hs[layer] * float(np.sum(ux * ux + uy * uy))
        Z += 0.5 * d * d * Hs[layer] * float(np.sum(q[1:-1, 1:-1] ** 2))
    dp = psi[1:-1, 1:-1, 0, 0] - psi[1:-1, 1:-1, 1, 0]
    E += 0.5 * d * d * m.H_1 * a * float(np.sum(dp * dp))
    return E, Z


def standard_model(M, P=None, Lx=4000.0 * KM, Ly=None, dt=60.0 * MINUTES, T=1.0 * DAY,
                   initial_kick=1e-6, U=0.1):
    """The parameter block of src/benchmarking/benchmarking.jl:6-26."""
    P = M if P is None else P
    Ly = Lx * P / M if Ly is None else Ly
    return make_model(1.0 * KM, 2.0 * KM, 2e-11, Lx, Ly, dt, T, U, M, P, Lx / M, 100.0, 1e-7,
                      40.0 * KM, initial_kick)
