"""A known answer for the WHOLE step of the oracle that does not come from the restatement itself: a barotropic
Rossby wave.  (The reference's own tests pin operators one at a time - src/test.jl - and it holds no golden
trajectory; this pins the time stepper, the sign of the beta term, the biharmonic viscosity and the inversion
together, against the equations the reference discretises, src/model.jl:139-199.)

With equal layer depths, no mean flow and no bottom friction, psi_1 = psi_2 = A cos(kx + ly) is an exact solution of
the reference's semi-discrete equations: q_l = Lap(psi_l) + S (psi_other - psi_l) = -lam psi with lam the eigenvalue
of the 5-point Laplacian, the Arakawa Jacobian J(q, psi) = -lam J(psi, psi) vanishes identically, and

    d/dt psi^ = (-nu lam + i beta kap / lam) psi^,    kap = sin(k dx) / dx   (centred x-difference),

i.e. psi(t) = A exp(-nu lam t) cos(kx + ly + beta kap t / lam): a westward-propagating, viscously decaying wave.
The reference's pinned node only shifts psi by a constant, removed before comparing.  Time integration is the only
approximation, and its error is dominated by the two Euler start-up steps of the reference's scheme
(src/model.jl:161; local error O(dt^2) each, carried along by AB3): halving dt must shrink the error 4x.
Measured: 4.07e-4, 1.02e-4, 2.54e-5 of the amplitude at dt = 60, 30, 15 min over 40 h (ratios 4.002, 4.001),
identically for the SuperLU and the spectral inversion."""
import numpy as np
import pytest

import qg_oracle as o


def wave_model(M, P, dt):
    L = 4.0e6
    return o.make_model(1000.0, 1000.0, 2e-11, L, L * P / M, dt, 86400.0, 0.0, M, P, L / M, 20000.0, 0.0, 4.0e4, 1e-6)


def wave(m, kx, ly, A, t):
    dx = m.dx
    k, l = 2 * np.pi * kx / m.Lx, 2 * np.pi * ly / m.Ly
    lam = (4.0 / dx ** 2) * (np.sin(0.5 * k * dx) ** 2 + np.sin(0.5 * l * dx) ** 2)
    kap = np.sin(k * dx) / dx
    x = (np.arange(m.M + 2) - 1) * dx          # index 1 <-> x = 0 (ghost ring included)
    y = (np.arange(m.P + 2) - 1) * dx
    ph = k * x[:, None] + l * y[None, :] + m.beta * kap / lam * t
    return A * np.exp(-m.visc * lam * t) * np.cos(ph), lam


def run_wave(M, P, dt, nsteps, backend):
    m = wave_model(M, P, dt)
    A = 1.0e4
    p0, lam = wave(m, 2, 1, A, 0.0)
    zeta = np.zeros((M + 2, P + 2, 2, 3), order="F")
    psi = np.zeros_like(zeta)
    for layer in range(2):
        psi[:, :, layer, 0] = p0
        zeta[:, :, layer, 0] = o.laplace_5p(np.asfortranarray(p0), m.dx)     # S (psi_2 - psi_1) = 0
    # the initial q is an eigenfunction of the discrete Laplacian
    assert np.allclose(zeta[1:-1, 1:-1, 0, 0], -lam * p0[1:-1, 1:-1], rtol=0, atol=1e-12 * lam * A)
    f = np.zeros_like(zeta)
    o.run_steps(m, zeta, psi, f, o.make_factors(m, backend), 1, nsteps)
    exact, _ = wave(m, 2, 1, A, nsteps * dt)
    errs = []
    for layer in range(2):
        got = psi[1:-1, 1:-1, layer, 0]
        got = got - got[0, 0]                                   # the pinned node: psi is defined up to a constant
        want = exact[1:-1, 1:-1] - exact[1, 1]
        errs.append(np.abs(got - want).max() / A)
    # q stays -lam psi (the Jacobian of a single wave vanishes, the layers stay equal)
    assert np.allclose(zeta[1:-1, 1:-1, 0, 0], zeta[1:-1, 1:-1, 1, 0], rtol=0, atol=1e-12 * lam * A)
    return max(errs), m, lam


@pytest.mark.parametrize("backend", ["direct", "spectral"])
def test_barotropic_rossby_wave_propagates_and_decays_as_the_equations_say(backend):
    M, P = 32, 24
    T = 40 * 3600.0
    e1, m, lam = run_wave(M, P, 3600.0, 40, backend)
    e2, _, _ = run_wave(M, P, 1800.0, 80, backend)
    # the wave has really moved and decayed over T, so the comparison is not vacuous
    k = 2 * np.pi * 2 / m.Lx
    phase = m.beta * np.sin(k * m.dx) / m.dx / lam * T
    decay = 1.0 - np.exp(-m.visc * lam * T)
    assert phase > 0.3 and decay > 0.03
    assert e1 < 6e-4 and e2 < 1.5e-4, (e1, e2)
    assert 3.8 < e1 / e2 < 4.2, (e1, e2)


def test_wave_solution_through_the_c_restatement():
    """The same known answer through oracle/qg_oracle.c (the CPU baseline and large-size checker)."""
    import qg_oracle_c as oc
    M, P, dt, n = 64, 32, 1800.0, 60
    m = wave_model(M, P, dt)
    A = 1.0e4
    p0, lam = wave(m, 2, 1, A, 0.0)
    zeta = np.zeros((M + 2, P + 2, 2, 3), order="F")
    psi = np.zeros_like(zeta)
    for layer in range(2):
        psi[:, :, layer, 0] = p0
        zeta[:, :, layer, 0] = o.laplace_5p(np.asfortranarray(p0), m.dx)
    f = np.zeros_like(zeta)
    oc.run_steps(m, zeta, psi, f, 1, n, 2)
    exact, _ = wave(m, 2, 1, A, n * dt)
    got = psi[1:-1, 1:-1, 0, 0] - psi[1, 1, 0, 0]
    want = exact[1:-1, 1:-1] - exact[1, 1]
    assert np.abs(got - want).max() / A < 2e-4
