"""The reference's own known-answer tests (reference src/test.jl), restated against the CPU
oracle.  These are what pins oracle/qg_oracle.py to the reference; the reference holds no
golden trajectory, so multi-step parity rests on the restatement itself (see the oracle
header)."""
import numpy as np
import pytest

import qg_oracle as o


def inflate(f, xs, ys):
    """src/schemes/laplacian.jl:94 / src/test.jl:106"""
    return np.asfortranarray(np.array([[f(x, y) for y in ys] for x in xs], dtype=np.float64))


def ref_model(U=2.0):
    # src/test.jl:9-23
    return o.make_model(1.0 * o.KM, 2.0 * o.KM, 2e-11, 4000.0 * o.KM, 4000.0 * o.KM, 15.0 * o.MINUTES,
                        0.5 * o.YEAR, U, 128, 128, 4000.0 * o.KM / 128, 100.0, 1e-7, 40.0 * o.KM, 1e-2)


def test_parameter_values_exact():
    """src/test.jl:8-44 — exact == on every derived parameter."""
    m = ref_model()
    expected_ratio = 0.5 * (1000 + 2000) / (40000 ** 2 * (1 / 1000 + 1 / 2000))
    assert expected_ratio == o.ratio_term(m)
    expected_S1 = 2 * expected_ratio / (1000 * 3000)
    assert expected_S1 == o.S1_plus(m)
    expected_S2 = 2 * expected_ratio / (2000 * 3000)
    assert expected_S2 == o.S2_minus(m)
    assert m.beta + expected_S1 * m.U == o.beta_1(m)
    assert m.beta - expected_S2 * m.U == o.beta_2(m)
    assert o.S_eig(m) == -1 / m.R_d ** 2
    assert (-o.S1_plus(m) - o.S2_minus(m)) == o.S_eig(m)


def test_explicit_laplacian_exact_on_cubic():
    """src/test.jl:55-69 — laplace_5p(x^3 + y^2, dx=1) == 6x + 2 exactly on a 10x10 array."""
    xs = ys = range(1, 11)
    u = inflate(lambda x, y: x ** 3 + y ** 2, xs, ys)
    true_lap = inflate(lambda x, y: 6 * x + 2, xs, ys)
    o.update_doubly_periodic_bc(true_lap)
    lap = o.laplace_5p(u, 1.0)
    assert np.array_equal(lap, true_lap)


def test_arakawa_second_order_and_sign():
    """src/test.jl:71-103 (the reference only prints the slope; the notebook
    scheme_validation.ipynb shows -2) — also pins the sign convention J(zeta, psi)."""
    Lx = Ly = 10.0
    A = lambda x, y: np.sin(2 * np.pi * x / Lx) * np.sin(2 * np.pi * y / Ly)
    B = lambda x, y: np.cos(2 * np.pi * x / Lx) * np.cos(2 * np.pi * y / Ly)
    Jt = lambda x, y: (-4 * np.pi ** 2 / (Lx * Ly) * np.cos(2 * np.pi * x / Lx) ** 2 * np.sin(2 * np.pi * y / Ly) ** 2
                       + 4 * np.pi ** 2 / (Lx * Ly) * np.sin(2 * np.pi * x / Lx) ** 2 * np.cos(2 * np.pi * y / Ly) ** 2)
    Ms = [8, 16, 32, 64, 128, 256]
    errs = []
    for M in Ms:
        dx = Lx / M
        xs = np.linspace(-dx, Lx, M + 2)
        ys = np.linspace(-dx, Ly, M + 2)
        err = o.J(dx, inflate(A, xs, ys), inflate(B, xs, ys)) - inflate(Jt, xs, ys)
        errs.append(dx * np.linalg.norm(err))
    slope = np.polyfit(np.log(Ms), np.log(errs), 1)[0]
    assert 1.7 < -slope < 2.3


@pytest.mark.parametrize("backend", ["direct", "spectral"])
@pytest.mark.parametrize("alpha", [0.0, -3.0])
def test_doubly_periodic_solves_converge(backend, alpha):
    """src/test.jl:105-148 (alpha = 0) and :150-193 (alpha = -3): slope in (1.7, 2.3).
    For alpha = 0 the pinned Poisson factor is used (u_true vanishes at the pinned node)."""
    x0, x1 = 0.0, 3.0
    L = x1 - x0
    u = lambda x, y: np.sin(2 * np.pi * x / L) * np.cos(2 * np.pi * y / L)
    f = lambda x, y: -(np.pi ** 2) * (u(x, y) * (4 / L ** 2 + 4 / L ** 2)) + alpha * u(x, y)
    Ms = [4, 8, 16, 32, 64]
    errs = []
    for M in Ms:
        dx = L / M
        xs = np.linspace(x0 - dx, x1, M + 2)
        ys = np.linspace(x0 - dx, x1, M + 2)
        b = inflate(f, xs, ys)
        if alpha == 0.0:
            fac = o.get_poisson_cholesky(M, M, dx) if backend == "direct" else o.get_poisson_spectral(M, M, dx)
            un = o.sp_solve_poisson(M, M, dx, b, factor=fac)
        else:
            fac = (o.get_helmholtz_cholesky(M, M, dx, alpha) if backend == "direct"
                   else o.get_helmholtz_spectral(M, M, dx, alpha))
            un = o.sp_solve_modified_helmholtz(M, M, dx, b, alpha, factor=fac)
        errs.append(dx * np.linalg.norm(un - inflate(u, xs, ys)))
    slope = np.polyfit(np.log(Ms), np.log(errs), 1)[0]
    assert 1.7 < -slope < 2.3


def test_P_times_P_inv_is_identity():
    """src/test.jl:195-217 — the *correct* pair P_matrix(H_1, H_2) * P_inv == I exactly."""
    m = ref_model()
    assert np.array_equal(o.P_matrix(m.H_1, m.H_2) @ o.P_inv_matrix(m), np.eye(2))


def test_evolve_psi_uses_H1_H1_projection():
    """evolve_psi! back-projects with P_matrix(H_1, H_1) = [1 -1; 1 1]
    (src/model.jl:173); the oracle reproduces the reference as it runs."""
    m = o.standard_model(8, 8)
    zeta, psi = o.initialise_model(m, seed=3)
    pf, hf = o.make_factors(m, "direct")
    o.evolve_psi(m, zeta, psi, pf, hf)
    Pinv = o.P_inv_matrix(m)
    zt = [Pinv[i, 0] * zeta[:, :, 0, 0] + Pinv[i, 1] * zeta[:, :, 1, 0] for i in range(2)]
    b = -zt[0][1:-1, 1:-1].ravel(order="F"); b[0] = 0
    t1 = o.add_doubly_periodic_boundaries(pf.solve(b).reshape((8, 8), order="F"))
    t2 = o.add_doubly_periodic_boundaries(hf.solve(-zt[1][1:-1, 1:-1].ravel(order="F")).reshape((8, 8), order="F"))
    assert np.array_equal(psi[:, :, 0, 0], 1.0 * t1 + (-1.0) * t2)
    assert np.array_equal(psi[:, :, 1, 0], 1.0 * t1 + 1.0 * t2)
    assert t1[1, 1] == 0.0   # pinned unknown


def test_periodic_1d_laplacian_matrix():
    """src/test.jl:229-238"""
    A = o.laplacian_1d_periodic(4).toarray()
    assert np.array_equal(A, np.array([[-2, 1, 0, 1], [1, -2, 1, 0], [0, 1, -2, 1], [1, 0, 1, -2.0]]))
    lam = np.sort(np.linalg.eigvalsh(A))
    assert np.allclose(lam, np.sort(2 * np.cos(2 * np.pi * np.arange(4) / 4) - 2))


def test_pinned_matrix_structure():
    """src/test.jl:246-276 — symmetric positive definite after pinning (4x4 and 10x5)."""
    for M, P in [(4, 4), (10, 5)]:
        A = (-o.construct_spA(M, P, 1.0, 0.0)).tolil()
        A[:, 0] = 0; A[0, :] = 0; A[0, 0] = 1
        A = A.toarray()
        assert np.array_equal(A, A.T)
        assert np.linalg.eigvalsh(A).min() > 0
        H = (-o.construct_spA(M, P, 1.0, -3.0)).toarray()
        assert np.linalg.eigvalsh(H).min() > 0


@pytest.mark.parametrize("M,P", [(8, 8), (16, 8), (24, 40), (33, 17)])
def test_direct_and_spectral_backends_agree(M, P):
    m = o.standard_model(M, P)
    zd, pd = o.run_model_no_output(m, seed=2, backend="direct", total_steps=10)
    zs, ps = o.run_model_no_output(m, seed=2, backend="spectral", total_steps=10)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(ps, pd) < 1e-12 and rel(zs, zd) < 1e-13


def test_pinned_solve_with_nonzero_mean_rhs():
    """The spectral statement of the pin (drop the mean mode, dump the residual into node 0,
    shift) equals the reference's pinned matrix for an arbitrary right-hand side."""
    rng = np.random.default_rng(0)
    b = rng.random(12 * 10); b[0] = 0
    d = o.get_poisson_cholesky(12, 10, 0.7).solve(b)
    s = o.get_poisson_spectral(12, 10, 0.7).solve(b)
    assert np.abs(d - s).max() / np.abs(d).max() < 1e-12


@pytest.mark.parametrize("name", ["traj_8x8_s10", "traj_16x8_s10", "traj_24x40_s10", "traj_64x64_s10"])
def test_oracle_reproduces_golden(name, golden_dir):
    g = np.load(f"{golden_dir}/{name}.npz")
    m = o.standard_model(int(g["M"]), int(g["P"]), dt=float(g["dt"]), initial_kick=float(g["kick"]))
    zeta, psi = o.initialise_model(m, seed=int(g["seed"]))
    f = np.zeros_like(zeta)
    o.run_steps(m, zeta, psi, f, o.make_factors(m, "spectral"), 1, int(g["steps"]))
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(psi, g["psi"]) < 1e-12
    assert rel(zeta, g["zeta"]) < 1e-13
    assert rel(f, g["f_store"]) < 1e-11
    E, Z = o.diagnostics(m, zeta, psi)
    assert abs(E - g["E"]) / g["E"] < 1e-12 and abs(Z - g["Z"]) / g["Z"] < 1e-12


def test_initial_condition_layout_and_assert():
    m = o.standard_model(6, 5)
    zeta, psi = o.initialise_model(m, seed=1)
    assert zeta.shape == (8, 7, 2, 3) and zeta.flags.f_contiguous
    assert np.all(zeta[:, :, :, 1:] == 0) and np.all(psi[:, :, :, 1:] == 0)
    for a in (zeta[:, :, 0, 0], psi[:, :, 1, 0]):
        assert np.array_equal(a[0, 1:-1], a[-2, 1:-1]) and np.array_equal(a[1:-1, -1], a[1:-1, 1])
        assert a[0, 0] == a[-2, -2] and a[-1, 0] == a[1, -2]
    bad = o.make_model(1000.0, 2000.0, 2e-11, 4e6, 4e6, 60.0, 600.0, 0.0, 8, 8, 5e5, 100.0, 1e-7, 4e4, 1e-6)
    with pytest.raises(AssertionError):
        o.initialise_model(bad)   # sign(beta_1) == -sign(beta_2) fails, src/model.jl:38
