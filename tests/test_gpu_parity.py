"""GPU parity tests: the CUDA path, called through the C ABI (qgb200.Session is a thin ctypes
wrapper), against the CPU oracle on identical seeded inputs, against the committed golden
trajectories, and through size-independent properties at the full benchmark sizes.

Tolerances are the ones BASELINE.json's north_star states: psi and q <= 1e-10 relative
(max norm over the field) after 10 steps; energy and enstrophy <= 1e-8 relative after
1000 steps.  Bit-exactness is not expected (different solver, FMA contraction)."""
import numpy as np
import pytest

import qg_oracle as o
import qg_oracle_c as oc
import qgb200

pytestmark = pytest.mark.gpu

TOL_FIELD = 1e-10
TOL_DIAG = 1e-8


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def models(M, P, **kw):
    mo = o.standard_model(M, P, **kw)
    mg = qgb200.BaroclinicModel(mo.H_1, mo.H_2, mo.beta, mo.Lx, mo.Ly, mo.dt, mo.T, mo.U, mo.M, mo.P, mo.dx,
                                mo.visc, mo.r, mo.R_d, mo.initial_kick)
    return mo, mg


def gpu_run(mg, zeta, psi, f, first, nsteps):
    z, p, ff = zeta.copy(order="F"), psi.copy(order="F"), f.copy(order="F")
    with qgb200.Session(mg) as s:
        s.upload(z, p, ff)
        s.step(first, nsteps)
        s.download(z, p, ff)
        E, Z = s.diagnostics()
        n = s.launch_count()
    assert n >= 4 * nsteps   # K1, K2, K3 (k = 0 column included), K4 per step
    return z, p, ff, E, Z


@pytest.mark.parametrize("name", ["traj_8x8_s10", "traj_16x8_s10", "traj_24x40_s10", "traj_64x64_s10"])
def test_ten_steps_against_golden(name, golden_dir):
    g = np.load(f"{golden_dir}/{name}.npz")
    mo, mg = models(int(g["M"]), int(g["P"]), dt=float(g["dt"]), initial_kick=float(g["kick"]))
    zeta, psi = o.initialise_model(mo, seed=int(g["seed"]))
    z, p, f, E, Z = gpu_run(mg, zeta, psi, np.zeros_like(zeta), 1, int(g["steps"]))
    for lvl in range(3):   # all three history levels
        assert rel(p[:, :, :, lvl], g["psi"][:, :, :, lvl]) < TOL_FIELD
        assert rel(z[:, :, :, lvl], g["zeta"][:, :, :, lvl]) < TOL_FIELD
        assert rel(f[:, :, :, lvl], g["f_store"][:, :, :, lvl]) < TOL_FIELD
    assert abs(E - g["E"]) / g["E"] < TOL_DIAG and abs(Z - g["Z"]) / g["Z"] < TOL_DIAG


@pytest.mark.parametrize("M,P,backend", [
    (8, 8, "direct"), (9, 7, "direct"), (3, 3, "direct"), (4, 5, "direct"), (16, 33, "direct"),
    (40, 24, "direct"), (56, 56, "direct"), (128, 128, "direct"), (128, 100, "spectral"),
    (512, 512, "spectral"), (256, 1100, "spectral"), (1024, 1024, "spectral"), (2048, 96, "spectral"),
    (8192, 64, "spectral"), (64, 4160, "spectral"), (16384, 64, "spectral"),
    # persistent y-solve with a ragged last CTA: 33 chunks over 4 CTAs of 9 (two 144-row TMA boxes
    # per tile), 65 chunks over 8 CTAs of 9 (the last CTA owns 2 chunks), and radix-16 rows (M = 256)
    (256, 1056, "spectral"), (64, 2080, "spectral"), (256, 256, "direct"),
    # one chunk per column (the smoke-test shape) and 17 chunks over 2 CTAs of 9
    (64, 32, "direct"), (128, 544, "spectral"),
    # more than 4096 rows: clusters of 16 CTAs (non-portable size), k = 0 column from k3_pre
    (64, 8192, "spectral"), (128, 6144, "spectral"),
    # M = 4096: ring-buffered TMA-fed row transforms; 5 rows over 2 CTAs (one group without a row),
    # 96 rows (one row per group), 300 rows over 148 CTAs (2 or 3 rows per CTA: the ring wraps once)
    (4096, 5, "spectral"), (4096, 96, "spectral"), (4096, 300, "spectral"),
    # M = 16384 with enough rows for several (row, field) units per CTA
    (16384, 160, "spectral"),
])
def test_ten_steps_against_oracle(M, P, backend):
    """psi, q (and the RHS history) after 10 steps; covers power-of-two and general M,
    ragged y chunks, P beyond one cluster's register capacity, odd sizes."""
    mo, mg = models(M, P)
    zeta, psi = o.initialise_model(mo, seed=1)
    f = np.zeros_like(zeta)
    z, p, ff, _, _ = gpu_run(mg, zeta, psi, f, 1, 10)
    o.run_steps(mo, zeta, psi, f, o.make_factors(mo, backend), 1, 10)
    assert rel(p, psi) < TOL_FIELD
    assert rel(z, zeta) < TOL_FIELD
    assert rel(ff, f) < TOL_FIELD
    for l in range(2):   # ghost cells are periodic images
        a = p[:, :, l, 0]
        assert np.array_equal(a[0, 1:-1], a[-2, 1:-1]) and np.array_equal(a[1:-1, -1], a[1:-1, 1])
        assert a[0, 0] == a[-2, -2] and a[-1, 0] == a[1, -2] and a[0, -1] == a[-2, 1] and a[-1, -1] == a[1, 1]


def test_long_row_cluster_pair_transforms_match_oracle():
    """The alternative M = 16384 row transforms (one transform per cluster of two CTAs, DSMEM exchange;
    QG_FFT_PAIR=1, not the default: measured slower) against the oracle, in a subprocess because the
    switch is read once per process.  320 units over 148 clusters: two or three per cluster, so the
    staging hand-shake and both mbarrier phases are exercised."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import qg_oracle as o, test_gpu_parity as t\n"
        "mo, mg = t.models(16384, 160)\n"
        "zeta, psi = o.initialise_model(mo, seed=1); f = np.zeros_like(zeta)\n"
        "z, p, ff, _, _ = t.gpu_run(mg, zeta, psi, f, 1, 10)\n"
        "o.run_steps(mo, zeta, psi, f, o.make_factors(mo, 'spectral'), 1, 10)\n"
        "print('PAIR', t.rel(p, psi), t.rel(z, zeta))\n"
        "assert t.rel(p, psi) < 1e-10 and t.rel(z, zeta) < 1e-10\n" % [p for p in sys.path if p])
    env = dict(os.environ, QG_FFT_PAIR="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "PAIR" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_reference_call_pattern_evolve_zeta_and_psi():
    """The reference's own loop (src/run_model_no_output.jl:3-16) written with the mirrored
    function API on host arrays, step by step, Euler (1, 2) and AB3 (3+) branches."""
    mo, mg = models(32, 48)
    zeta, psi = qgb200.initialise_model(mg, seed=4)
    zo, po = o.initialise_model(mo, seed=4)
    assert np.array_equal(zeta, zo) and np.array_equal(psi, po)
    pc = qgb200.get_poisson_cholesky(mg.M, mg.P, mg.dx)
    hc = qgb200.get_helmholtz_cholesky(mg.M, mg.P, mg.dx, qgb200.S_eig(mg))
    f = np.zeros((mg.M + 2, mg.P + 2, 2, 3), order="F")
    fo = f.copy(order="F")
    fac = o.make_factors(mo, "direct")
    for t in range(1, 6):
        qgb200.evolve_zeta(mg, zeta, psi, t, f)
        o.evolve_zeta(mo, zo, po, t, fo)
        assert rel(zeta, zo) < 1e-12 and rel(f, fo) < 1e-12
        qgb200.evolve_psi(mg, zeta, psi, pc, hc)
        o.evolve_psi(mo, zo, po, *fac)
        assert rel(psi, po) < 1e-11
    qgb200.close_sessions()


def test_evolve_psi_on_nonzero_mean_pv_matches_pinned_matrix():
    """Arbitrary (non zero-mean) q exercises the pinned node exactly as the reference's
    matrix does (src/schemes/laplacian.jl:66-75, src/model.jl:185)."""
    mo, mg = models(20, 12)
    rng = np.random.default_rng(7)
    zeta = np.zeros((22, 14, 2, 3), order="F")
    for l in range(2):
        zeta[:, :, l, 0] = o.update_doubly_periodic_bc(np.asfortranarray(rng.random((22, 14)) + 0.5))
    psi = np.zeros_like(zeta)
    zo, po = zeta.copy(order="F"), psi.copy(order="F")
    qgb200.evolve_psi(mg, zeta, psi, qgb200.get_poisson_cholesky(20, 12, mg.dx),
                      qgb200.get_helmholtz_cholesky(20, 12, mg.dx, qgb200.S_eig(mg)))
    o.evolve_psi(mo, zo, po, *o.make_factors(mo, "direct"))
    assert rel(psi, po) < 1e-11
    qgb200.close_sessions()


def test_run_model_no_output_config1():
    """BASELINE.json config 1: the reference's benchmark block (src/benchmarking/benchmarking.jl:6-26),
    M = P = 128, dt = 60 min, T = 1 day -> 24 steps, through the mirrored driver."""
    mo, mg = models(128, 128)
    r = o.seeded_random_fields(mo, 1)
    zeta, psi = qgb200.run_model_no_output(mg, rand_fields=r)
    zo, po = o.run_model_no_output(mo, seed=1, backend="direct")
    assert rel(psi, po) < TOL_FIELD and rel(zeta, zo) < TOL_FIELD


def test_energy_enstrophy_after_1000_steps(golden_dir):
    rows = np.load(f"{golden_dir}/diag_1000steps.npy")
    for M, P, dt, kick, steps, E0, Z0 in rows:
        mo, mg = models(int(M), int(P), dt=float(dt), initial_kick=float(kick))
        zeta, psi = o.initialise_model(mo, seed=1)
        _, _, _, E, Z = gpu_run(mg, zeta, psi, np.zeros_like(zeta), 1, int(steps))
        assert abs(E - E0) / E0 < TOL_DIAG, (M, P, E, E0)
        assert abs(Z - Z0) / Z0 < TOL_DIAG, (M, P, Z, Z0)


def test_ensemble_members_are_independent_runs():
    """Members share parameters, differ in initial condition; each equals its own solo run."""
    mo, mg = models(64, 64)
    nm = 5
    zs, ps = [], []
    for m in range(nm):
        z, p = o.initialise_model(mo, seed=1 + m)
        zs.append(z); ps.append(p)
    Z = np.asfortranarray(np.stack(zs, axis=-1)); Pm = np.asfortranarray(np.stack(ps, axis=-1))
    F = np.zeros_like(Z)
    with qgb200.Session(mg, members=nm) as s:
        s.upload(Z, Pm, F)
        s.step(1, 6)
        s.download(Z, Pm, F)
        E, Zs = s.diagnostics()
    fac = o.make_factors(mo, "spectral")
    for m in range(nm):
        f = np.zeros_like(zs[m])
        o.run_steps(mo, zs[m], ps[m], f, fac, 1, 6)
        assert rel(Pm[..., m], ps[m]) < TOL_FIELD and rel(Z[..., m], zs[m]) < TOL_FIELD
        Eo, Zo = o.diagnostics(mo, zs[m], ps[m])
        assert abs(E[m] - Eo) / Eo < 1e-12 and abs(Zs[m] - Zo) / Zo < 1e-12


def test_single_use_solves_converge_second_order():
    """src/test.jl:105-193 run end to end on the CUDA solver (sp_solve_poisson for alpha = 0,
    sp_solve_modified_helmholtz for alpha = -3)."""
    L = 3.0
    u = lambda x, y: np.sin(2 * np.pi * x / L) * np.cos(2 * np.pi * y / L)
    for alpha in (0.0, -3.0):
        f = lambda x, y: -(np.pi ** 2) * (u(x, y) * (8 / L ** 2)) + alpha * u(x, y)
        Ms, errs = [4, 8, 16, 32, 64], []
        for M in Ms:
            dx = L / M
            xs = np.linspace(-dx, L, M + 2)
            b = np.asfortranarray(np.array([[f(x, y) for y in xs] for x in xs]))
            ut = np.array([[u(x, y) for y in xs] for x in xs])
            un = qgb200.sp_solve_poisson(M, M, dx, b) if alpha == 0.0 else \
                qgb200.sp_solve_modified_helmholtz(M, M, dx, b, alpha)
            ref = o.sp_solve_poisson(M, M, dx, b) if alpha == 0.0 else o.sp_solve_modified_helmholtz(M, M, dx, b, alpha)
            assert np.abs(un - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
            errs.append(dx * np.linalg.norm(un - ut))
        slope = np.polyfit(np.log(Ms), np.log(errs), 1)[0]
        assert 1.7 < -slope < 2.3


def test_headline_grid_properties_4096():
    """Config 3 (4096 x 4096, dt = 5 min): too large for the direct oracle, so check
    size-independent properties of one full step on the device result:
      - the inversion residual: Lap(psi~1) = q~1 away from the pinned node and
        (Lap + S_eig) psi~2 = q~2 everywhere, recomputed in NumPy from the downloaded fields;
      - the pinned unknown psi~1(0,0) is zero;
      - PV is conserved by the flux-form RHS: sum(q+ - q) / (sum |q| ) is round-off for layer 1
        (layer 2 has the r*Lap(psi) sink whose sum is also zero on a periodic grid);
      - against the oracle (C restatement, itself pinned to the NumPy one in tests/test_oracle_c.py) for
        the 10 steps north_star names (2 Euler + 8 AB3): psi, q and the RHS history to <= 1e-10."""
    M = P = 4096
    mo, mg = models(M, P, dt=300.0)
    zeta, psi = o.initialise_model(mo, seed=1)
    f = np.zeros_like(zeta)
    z, p, ff, E, Z = gpu_run(mg, zeta, psi, f, 1, 10)
    Pinv = o.P_inv_matrix(mo)
    q = z[1:-1, 1:-1, :, 0]
    qt = [Pinv[i, 0] * q[:, :, 0] + Pinv[i, 1] * q[:, :, 1] for i in range(2)]
    # undo the (H1,H1) back-projection: psi1 = t1 - t2, psi2 = t1 + t2
    t1 = 0.5 * (p[:, :, 0, 0] + p[:, :, 1, 0])
    t2 = 0.5 * (p[:, :, 1, 0] - p[:, :, 0, 0])
    idx2 = (1.0 / mo.dx) ** 2
    lap = lambda u: (u[:-2, 1:-1] + u[2:, 1:-1] - 4 * u[1:-1, 1:-1] + u[1:-1, :-2] + u[1:-1, 2:]) * idx2
    r1 = lap(t1) - qt[0]
    r1[0, 0] = 0.0
    r2 = lap(t2) + o.S_eig(mo) * t2[1:-1, 1:-1] - qt[1]
    assert np.abs(r1).max() / np.abs(qt[0]).max() < 1e-9
    assert np.abs(r2).max() / np.abs(qt[1]).max() < 1e-9
    assert abs(t1[1, 1]) <= 1e-12 * np.abs(t1).max()
    for l in range(2):
        dq = (z[1:-1, 1:-1, l, 0] - z[1:-1, 1:-1, l, 1]).sum()
        assert abs(dq) / np.abs(z[1:-1, 1:-1, l, 0]).sum() < 1e-11
    oc.run_steps(mo, zeta, psi, f, 1, 10)
    for lvl in range(3):
        assert rel(p[:, :, :, lvl], psi[:, :, :, lvl]) < TOL_FIELD
        assert rel(z[:, :, :, lvl], zeta[:, :, :, lvl]) < TOL_FIELD
        assert rel(ff[:, :, :, lvl], f[:, :, :, lvl]) < TOL_FIELD
    Eo, Zo = o.diagnostics(mo, zeta, psi)
    assert abs(E - Eo) / Eo < TOL_DIAG and abs(Z - Zo) / Zo < TOL_DIAG


def test_energy_enstrophy_after_1000_steps_1024():
    """SURVEY.md 8(d) correctness gate at config 2's grid: domain-integrated energy and enstrophy after
    1000 steps of 1024 x 1024 (dt = 30 min) <= 1e-8 relative against the oracle (C restatement), and the
    fields themselves, which stay far inside the 10-step tolerance."""
    mo, mg = models(1024, 1024, dt=1800.0)
    zeta, psi = o.initialise_model(mo, seed=1)
    f = np.zeros_like(zeta)
    with qgb200.Session(mg) as s:
        s.upload_initial(zeta, psi)
        s.step(1, 1000)
        z, p = s.new_state_array(), s.new_state_array()
        s.download(zeta=z, psi=p)
        E, Z = s.diagnostics()
    oc.run_steps(mo, zeta, psi, f, 1, 1000)
    Eo, Zo = o.diagnostics(mo, zeta, psi)
    assert abs(E - Eo) / Eo < TOL_DIAG and abs(Z - Zo) / Zo < TOL_DIAG, (E, Eo, Z, Zo)
    assert rel(p[:, :, :, 0], psi[:, :, :, 0]) < 1e-9 and rel(z[:, :, :, 0], zeta[:, :, :, 0]) < 1e-9


def test_reupload_of_zeta_and_psi_keeps_the_rhs_history():
    """qg_upload_state(zeta, psi, NULL) on a handle that has stepped: f_store stays on the device
    ("NULL leaves that array untouched") and must still line up with the freshly uploaded zeta, so
    the next AB3 step reads the right two older right-hand sides."""
    mo, mg = models(40, 24)
    zeta, psi = o.initialise_model(mo, seed=5)
    f = np.zeros_like(zeta)
    fac = o.make_factors(mo, "direct")
    with qgb200.Session(mg) as s:
        s.upload(zeta, psi, f)
        s.step(1, 5)
        z, p = s.new_state_array(), s.new_state_array()
        s.download(zeta=z, psi=p)
        s.upload(z, p, None)            # zeta and psi only
        s.step(6, 3)
        s.download(zeta=z, psi=p)
    o.run_steps(mo, zeta, psi, f, fac, 1, 8)
    assert rel(p, psi) < TOL_FIELD and rel(z, zeta) < TOL_FIELD


def test_error_reporting():
    mo, mg = models(16, 16)
    with qgb200.Session(mg) as s:
        with pytest.raises(qgb200.QGError) as ei:
            s.step(1, 1)   # nothing uploaded yet
        assert ei.value.code == -5
        z = s.new_state_array()
        s.upload(z, z.copy(order="F"), None)
        with pytest.raises(qgb200.QGError):
            s.evolve_zeta(0)   # timestep is 1-based
        with pytest.raises(ValueError):
            s.upload(np.zeros((16, 16, 2, 3)), None, None)


def test_y_slab_decomposition_on_two_gpus():
    """One run split into y-slabs over 2 GPUs equals the oracle's global solution, with the per-step
    exchanges done by in-kernel NVLink peer stores + flag barriers and, as a cross-check, by NCCL
    (QG_DIST_NCCL=1); needs two visible GPUs, launched as one process per GPU."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for M, P, nccl in ((256, 256, False), (512, 1024, False), (256, 256, True)):
        env = dict(os.environ)
        if nccl:
            env["QG_DIST_NCCL"] = "1"
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29541",
                              os.path.join(root, "tests", "dist_slab_check.py"), str(M), str(P), "10"],
                             capture_output=True, text=True, timeout=600, env=env)
        assert "SLAB_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_initial_upload_and_graph_replay_match_plain_path(monkeypatch):
    """qg_upload_initial_state (level 1 only, history and f_store zeroed on the device) followed by
    a long qg_step (CUDA-graph replay of the 3-step cycle) gives bit-identical results to the
    full upload and step-by-step launches."""
    mo, mg = models(96, 64)
    zeta, psi = o.initialise_model(mo, seed=9)
    f = np.zeros_like(zeta)
    za, pa, fa = zeta.copy(order="F"), psi.copy(order="F"), f.copy(order="F")
    with qgb200.Session(mg) as s:
        s.upload_initial(za, pa)
        s.step(1, 23)                       # 2 Euler + 21 AB3 steps: 7 graph replays
        s.download(za, pa, fa)
    zb, pb, fb = zeta.copy(order="F"), psi.copy(order="F"), f.copy(order="F")
    with qgb200.Session(mg) as s:
        s.upload(zb, pb, fb)
        for t in range(1, 24):              # one step per call: never enough steps left for a graph
            s.step(t, 1)
        s.download(zb, pb, fb)
    assert np.array_equal(za, zb) and np.array_equal(pa, pb) and np.array_equal(fa, fb)
    o.run_steps(mo, zeta, psi, f, o.make_factors(mo, "spectral"), 1, 23)
    assert rel(pa, psi) < TOL_FIELD and rel(za, zeta) < TOL_FIELD


@pytest.mark.gpu
def test_run_model_snapshots_and_restart(tmp_path):
    """run_model (src/run_model.jl:55-93): the samples written every sample_timestep are the newest
    level of the trajectory at exactly those steps, the returned arrays equal run_model_no_output's,
    and a run resumed from a restart file continues bit for bit."""
    mo, mg = models(64, 64)
    fn = str(tmp_path / "run.npz")
    zeta, psi = qgb200.run_model(mg, fn, True, seed=3, sample_timestep=4, total_steps=10)
    meta, snaps = qgb200.load_run(fn)
    assert sorted(snaps) == [0, 4, 8] and meta["dt"] == mg.dt
    zo, po = o.initialise_model(mo, seed=3)
    fo = np.zeros_like(zo)
    fac = o.make_factors(mo, "direct")
    assert np.array_equal(snaps[0][0], zo[:, :, :, 0]) and np.array_equal(snaps[0][1], po[:, :, :, 0])
    done = 0
    for t in (4, 8, 10):
        o.run_steps(mo, zo, po, fo, fac, done + 1, t - done)
        done = t
        if t in snaps:
            assert rel(snaps[t][0], zo[:, :, :, 0]) < TOL_FIELD and rel(snaps[t][1], po[:, :, :, 0]) < TOL_FIELD
    assert rel(zeta, zo) < TOL_FIELD and rel(psi, po) < TOL_FIELD
    z2, p2 = qgb200.run_model_no_output(mg, seed=3, total_steps=10)
    assert np.array_equal(z2, zeta) and np.array_equal(p2, psi)
    # restart: 6 steps, save, resume 4 more == 10 steps straight
    zi, pi_ = qgb200.initialise_model(mg, seed=3)
    rf = str(tmp_path / "restart.npz")
    with qgb200.Session(mg) as s:
        s.upload_initial(zi, pi_)
        s.step(1, 6)
        qgb200.save_restart(rf, s, 6)
    z3, p3, f3, t3 = qgb200.resume_model(rf, 4)
    assert t3 == 10 and np.array_equal(z3, zeta) and np.array_equal(p3, psi)


@pytest.mark.gpu
def test_device_initial_condition_matches_host_initialise_model():
    """qg_init_state == initialise_model (src/model.jl:37-62) fed with the same uniform draws: psi bit
    for bit (Philox stream restated in tests/philox_ref.py), q to rounding (the device contracts the
    Laplacian's last multiply-add); history levels and f_store zero; members draw different noise."""
    import philox_ref
    mo, mg = models(48, 40)
    nm = 3
    with qgb200.Session(mg, members=nm) as s:
        s.init_state(12345)
        z, p, f = s.new_state_array(), s.new_state_array(), s.new_state_array()
        s.download(z, p, f)
    assert not f.any() and not z[:, :, :, 1:].any() and not p[:, :, :, 1:].any()
    for m in range(nm):
        zh, ph = qgb200.initialise_model(mg, rand_fields=philox_ref.rand_fields(48, 40, 12345, member=m))
        assert np.array_equal(p[..., m], ph)
        assert rel(z[..., m], zh) < 1e-14
    assert not np.array_equal(p[..., 0], p[..., 1])
    u = p[1:-1, 1:-1, :, 0, :] / (mg.initial_kick * mg.U * mg.Ly)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    # and the run it starts is the run the host IC starts
    za, pa = qgb200.run_model_no_output(mg, device_ic=7, total_steps=5)
    zb, pb = qgb200.run_model_no_output(mg, rand_fields=philox_ref.rand_fields(48, 40, 7), total_steps=5)
    assert rel(pa, pb) < TOL_FIELD and rel(za, zb) < TOL_FIELD


@pytest.mark.gpu
def test_two_devices_in_one_process():
    """Handles on two GPUs held by ONE process (launch attributes are per device): same input, same
    steps, bit-identical output; needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mo, mg = models(4096, 64)            # radix-16 rows, persistent y-solve
    zeta, psi = o.initialise_model(mo, seed=2)
    out = []
    for dev in (0, 1):
        z, p, f = zeta.copy(order="F"), psi.copy(order="F"), np.zeros_like(zeta)
        with qgb200.Session(mg, device=dev) as s:
            s.upload(z, p, f)
            s.step(1, 4)
            s.download(z, p, f)
        out.append((z, p, f))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_run_model_monitor_time_series_and_running_extrema(tmp_path):
    """the diagnostics suite of run_model - energy / enstrophy time series and the
    running extrema the reference's dead update_max / update_min (src/run_model.jl:41-53) were
    meant to keep - against the oracle's trajectory sampled at the same steps; extrema are exact
    functions of fields that agree to 1e-10."""
    mo, mg = models(96, 64)
    fn = str(tmp_path / "run.npz")
    mon = {}
    zeta, psi = qgb200.run_model(mg, fn, True, seed=3, sample_timestep=6, total_steps=12, monitor_every=4, monitor=mon)
    series, running = mon["monitor"], mon["monitor_running"]
    assert series.shape == (4, 3 + 8) and list(series[:, 0]) == [0, 4, 8, 12]
    zo, po = o.initialise_model(mo, seed=3)
    fo = np.zeros_like(zo)
    fac = o.make_factors(mo, "direct")
    done, run_ref = 0, None
    for row in series:
        t = int(row[0])
        o.run_steps(mo, zo, po, fo, fac, done + 1, t - done)
        done = t
        Eo, Zo = o.diagnostics(mo, zo, po)
        assert abs(row[1] - Eo) / Eo < 1e-11 and abs(row[2] - Zo) / Zo < 1e-11
        ex = []
        for arr in (zo, po):
            for l in range(2):
                a = arr[1:-1, 1:-1, l, 0]
                ex += [a.max(), a.min()]
        assert np.allclose(row[3:], ex, rtol=1e-10, atol=0.0)
        run_ref = ex if run_ref is None else [max(r, v) if k % 2 == 0 else min(r, v) for k, (r, v) in enumerate(zip(run_ref, ex))]
    assert np.allclose(running, run_ref, rtol=1e-10, atol=0.0)
    with np.load(fn) as z:
        assert np.array_equal(z["monitor"], series) and np.array_equal(z["monitor_running"], running)
    assert qgb200.update_max(1.0, 2.0) == 2.0 and qgb200.update_max(3.0, 2.0) == 3.0 and qgb200.update_min(1.0, 2.0) == 1.0
    # ensembles: one row of extrema per member
    with qgb200.Session(mg, members=2) as s:
        s.init_state(5)
        e = s.extrema()
        z, p = s.new_state_array(), s.new_state_array()
        s.download(zeta=z, psi=p)
    assert e.shape == (2, 8) and e[1, 4] == p[1:-1, 1:-1, 0, 0, 1].max() and e[0, 1] == z[1:-1, 1:-1, 0, 0, 0].min()


def test_repeated_runs_are_bit_identical_4096():
    """The same run repeated on one handle gives bit-identical fields.  Guards the tile-reuse race found
    in round 2 (k3_ysolve_pipe issued the next slab's TMA copies behind a barrier that directly followed
    the 32 shared-memory loads of the current tile: the barrier orders the loads' issue, not their
    completion, and about one 43-step run in ten at 4096 x 4096 came out different in the last bits;
    profiles/r02/determinism_k3_tile_race.log).  Every kernel of the step is deterministic by
    construction (fixed-order reductions), so any difference is a race."""
    import hashlib
    mo, mg = models(4096, 4096, dt=300.0)
    seen = set()
    with qgb200.Session(mg) as s:
        z1 = np.zeros((4098, 4098, 2), order="F")
        p1 = np.zeros((4098, 4098, 2), order="F")
        for rep in range(8):
            s.init_state(3)
            s.step(1, 40)
            s.snapshot_begin(z1, p1)
            s.snapshot_end()
            seen.add(hashlib.sha256(z1.tobytes() + p1.tobytes()).hexdigest())
    assert len(seen) == 1, f"{len(seen)} different results in 8 identical runs"
