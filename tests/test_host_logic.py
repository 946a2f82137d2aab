"""CPU-side checks of the host shim and of the C-ABI library: it loads, exports every symbol
include/qgb200.h declares, refuses to run without a GPU (no CPU fallback), and the host
mirror evaluates the reference's parameter algebra exactly."""
import ctypes
import os
import re

import numpy as np
import pytest

import qgb200
from qgb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def header_functions():
    src = open(os.path.join(ROOT, "include", "qgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    _ensure_built()
    names = header_functions()
    assert len(names) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libqgb200.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"
    assert qgb200.load().qg_abi_version() == 2


def test_peer_exchange_is_refused_when_two_ranks_share_a_gpu():
    """qg_dist_ipc_import's precondition, checked without a GPU: exports (3 IPC handles + the GPU's
    UUID at byte 192, 256 bytes per rank) from the same device must be refused - two flag-waiting
    barrier kernels on one GPU are not guaranteed to be co-resident (B200_PROFILING.md, Xid 109)."""
    _ensure_built()
    lib = qgb200.load()
    blob = lambda uuid: bytes(192) + bytes(uuid) + bytes(48)
    u = [bytes([r + 1] * 16) for r in range(8)]
    distinct = b"".join(blob(x) for x in u)
    assert lib.qg_dist_ipc_blobs_share_device(distinct, 8) == 0
    assert lib.qg_dist_ipc_blobs_share_device(distinct, 2) == 0
    for n, dup in ((2, (0, 1)), (4, (1, 3)), (8, (0, 7))):
        v = list(u[:n])
        v[dup[1]] = v[dup[0]]
        assert lib.qg_dist_ipc_blobs_share_device(b"".join(blob(x) for x in v), n) == 1
    assert lib.qg_dist_ipc_blobs_share_device(None, 2) == -1 and lib.qg_dist_ipc_blobs_share_device(distinct, 9) == -1
    assert qgb200.Session.IPC_BLOB == 256


def test_param_struct_matches_header_layout():
    # int32 M, P; 8 doubles; 2x double[4]; 3 doubles
    assert ctypes.sizeof(qgb200.qg_params) == 8 + 8 * 8 + 2 * 32 + 3 * 8
    assert qgb200.qg_params.dx.offset == 8 and qgb200.qg_params.Pinv.offset == 72


def test_no_gpu_means_loud_failure(have_gpu):
    _ensure_built()
    if have_gpu:
        pytest.skip("GPU present")
    m = qgb200.BaroclinicModel(1000., 2000., 2e-11, 4e6, 4e6, 3600., 86400., 0.1, 16, 16, 4e6 / 16, 100., 1e-7, 4e4, 1e-6)
    with pytest.raises(qgb200.QGError) as ei:
        qgb200.Session(m)
    assert ei.value.code == -4 and "no CPU fallback" in ei.value.message


def test_create_rejects_bad_arguments():
    _ensure_built()
    lib = qgb200.load()
    m = qgb200.BaroclinicModel(1000., 2000., 2e-11, 4e6, 4e6, 3600., 86400., 0.1, 2, 16, 4e6 / 16, 100., 1e-7, 4e4, 1e-6)
    p = qgb200.make_params(m)
    h = ctypes.c_void_p()
    assert lib.qg_create(ctypes.byref(p), 0, 1, None, ctypes.byref(h)) == -1
    assert b"M and P" in lib.qg_last_error(None)
    assert lib.qg_create(None, 0, 1, None, ctypes.byref(h)) == -1
    assert lib.qg_step(None, 1, 1) == -1 and lib.qg_destroy(None) == 0


def test_parameter_mirror_is_exact():
    """src/test.jl:8-44 on the product's host mirror."""
    m = qgb200.BaroclinicModel(1.0 * qgb200.KM, 2.0 * qgb200.KM, 2e-11, 4000.0 * qgb200.KM, 4000.0 * qgb200.KM,
                               15.0 * qgb200.MINUTES, 0.5 * qgb200.YEAR, 2.0, 128, 128, 4000.0 * qgb200.KM / 128,
                               100.0, 1e-7, 40.0 * qgb200.KM, 1e-2)
    ratio = 0.5 * (1000 + 2000) / (40000 ** 2 * (1 / 1000 + 1 / 2000))
    assert qgb200.ratio_term(m) == ratio
    assert qgb200.S1_plus(m) == 2 * ratio / (1000 * 3000)
    assert qgb200.S2_minus(m) == 2 * ratio / (2000 * 3000)
    assert qgb200.beta_1(m) == m.beta + qgb200.S1_plus(m) * m.U
    assert qgb200.beta_2(m) == m.beta - qgb200.S2_minus(m) * m.U
    assert qgb200.S_eig(m) == -1 / m.R_d ** 2 == (-qgb200.S1_plus(m) - qgb200.S2_minus(m))
    assert np.array_equal(qgb200.P_matrix(m.H_1, m.H_2) @ qgb200.P_inv_matrix(m), np.eye(2))
    assert m.H == 3000.0 and m.domain == qgb200.RectangularDomain(0.0, m.Lx, 0.0, m.Ly)
    p = qgb200.make_params(m)
    assert list(p.Pfwd) == [1.0, -1.0, 1.0, 1.0]          # P_matrix(H_1, H_1), src/model.jl:173
    assert list(p.Pinv) == list(qgb200.P_inv_matrix(m).ravel())
    assert (p.beta1, p.beta2, p.alpha) == (qgb200.beta_1(m), qgb200.beta_2(m), qgb200.S_eig(m))
    with pytest.raises(AttributeError):
        m.dt = 1.0


def test_host_mirror_matches_oracle_initial_condition():
    import qg_oracle as o
    mo = o.standard_model(12, 9)
    mg = qgb200.BaroclinicModel(mo.H_1, mo.H_2, mo.beta, mo.Lx, mo.Ly, mo.dt, mo.T, mo.U, mo.M, mo.P, mo.dx,
                                mo.visc, mo.r, mo.R_d, mo.initial_kick)
    zo, po = o.initialise_model(mo, seed=5)
    zg, pg = qgb200.initialise_model(mg, seed=5)
    assert np.array_equal(zo, zg) and np.array_equal(po, pg)
    bad = qgb200.BaroclinicModel(1000.0, 2000.0, 2e-11, 4e6, 4e6, 60.0, 600.0, 0.0, 8, 8, 5e5, 100.0, 1e-7, 4e4, 1e-6)
    with pytest.raises(AssertionError):
        qgb200.initialise_model(bad)


def test_plan_tokens_and_argument_checks():
    m = qgb200.BaroclinicModel(1000., 2000., 2e-11, 4e6, 4e6, 3600., 86400., 0.1, 16, 16, 4e6 / 16, 100., 1e-7, 4e4, 1e-6)
    pc = qgb200.get_poisson_cholesky(16, 16, m.dx)
    hc = qgb200.get_helmholtz_cholesky(16, 16, m.dx, qgb200.S_eig(m))
    assert pc.pinned and not hc.pinned
    z = np.zeros((18, 18, 2, 3), order="F")
    with pytest.raises(TypeError):
        qgb200.evolve_psi(m, z, z.copy(order="F"), hc, pc)   # swapped plans
    with pytest.raises(ValueError):
        qgb200.evolve_psi(m, z, z.copy(order="F"), pc, qgb200.get_helmholtz_cholesky(16, 16, m.dx, -1.0))


def test_run_model_output_container_roundtrip(tmp_path):
    """The run_model output file (reference keys zeta_<t>, psi_<t>, metadata; src/run_model.jl:6-20,
    70-73, 86-90) is appended to sample by sample and read back intact; no GPU involved."""
    from qgb200 import runs as rm
    m = qgb200.BaroclinicModel(1000.0, 2000.0, 2e-11, 4.0e6, 4.0e6, 3600.0, 10 * 86400.0, 0.1, 8, 8, 5.0e5, 100.0,
                               1e-7, 4.0e4, 1e-6)
    meta = rm.create_metadata(m)
    assert meta == {"dt": 3600.0, "T": 864000.0, "sample_interval": 86400.0, "sample_timestep": 24,
                    "total_steps": 240}
    f = str(tmp_path / "run.npz")
    rng = np.random.default_rng(0)
    snaps = {t: (rng.random((10, 10, 2)), rng.random((10, 10, 2))) for t in (0, 48, 96)}
    rm._append(f, zeta_0=snaps[0][0], psi_0=snaps[0][1],
               metadata=np.frombuffer(__import__("json").dumps(meta).encode(), dtype=np.uint8))
    for t in (48, 96):
        rm._append(f, **{f"zeta_{t}": np.asfortranarray(snaps[t][0]), f"psi_{t}": snaps[t][1]})
    meta2, got = rm.load_run(f)
    assert meta2 == meta and sorted(got) == [0, 48, 96]
    for t in snaps:
        assert np.array_equal(got[t][0], snaps[t][0]) and np.array_equal(got[t][1], snaps[t][1])
    lines = []
    rm.log_model_params(m, lines.append)
    assert lines[0] == "Parameters:" and any(l.startswith("Beta_1 = ") for l in lines) and lines[-1] == "Total steps = 240"


def test_philox_restatement_known_answer():
    """tests/philox_ref.py (the checker of qg_init_state's random stream) reproduces the published
    Philox4x32-10 known-answer vector for counter = key = 0 (Random123 kat_vectors)."""
    import philox_ref
    w = [int(x[0]) for x in philox_ref.philox_words(np.zeros(1, dtype=np.uint64), 0, 0)]
    assert w == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    u = philox_ref.philox_uniform(np.arange(200000, dtype=np.uint64), 1, 99)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3


def test_persistent_ysolve_algebra_matches_dense_cyclic_solve():
    """The formulation of k3_ysolve_pipe - 8-row segmented zero-carry sweeps, chunk -> CTA -> cluster
    affine closures, element-wise apply with the column-table coefficients cA / cB (tests/ysolve_model.py)
    - against a dense solve of the cyclic tridiagonal system of src/schemes/laplacian.jl:40-58 and
    against the recurrence form the first kernels use."""
    import ysolve_model as ym
    rng = np.random.default_rng(5)
    for e, n, nctas in ((0.3, 256, 2), (2.5e-6, 512, 4), (4.61, 128, 1), (1e-3, 1024, 8)):
        g = rng.standard_normal(n)
        d = -(2.0 + e)
        A = np.zeros((n, n))
        for j in range(n):
            A[j, j] = d
            A[j, (j - 1) % n] += 1.0
            A[j, (j + 1) % n] += 1.0
        ref = np.linalg.solve(A, g)
        u = ym.solve_cyclic_pipe(g, e, nctas)
        tol = 1e-12 / min(1.0, e)            # the system's own conditioning ~ 4/e
        assert np.abs(u - ref).max() <= tol * np.abs(ref).max()
        assert np.abs(u - ym.solve_cyclic(g, e)).max() <= 1e-13 / min(1.0, e) * np.abs(ref).max()
    # the coefficients: element-wise apply == carrying A and B through the two recurrences
    r = 0.9
    cA, cB = ym.column_table(r)
    b = rng.standard_normal(32)
    z, F, G = ym.chunk_double_sweep(b, r)
    Ac, Bc = 0.7, -1.3
    y = np.zeros(32); acc = Ac
    for i in range(32):
        acc = b[i] + r * acc; y[i] = acc
    zz = np.zeros(32); acc = Bc
    for i in range(31, -1, -1):
        acc = y[i] + r * acc; zz[i] = acc
    assert np.allclose(z + Ac * cA + Bc * cB, zz, rtol=1e-13, atol=1e-13)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference algorithm on the host cores, honouring --steps / --warmup, with the
    same `config` dictionary the GPU arm prints for that grid and GPU count, `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "64", "64",
                          "--steps", "3", "--warmup", "2", "--gpus", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["steps"] == 3 and line["warmup"] == 2 and line["n_gpus"] == 2
    assert line["config"] == bench.workload_config(64, 64, 2, 1, bench.model_args(64, 64)["dt"])
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_plan_probe_matches_the_model_of_the_y_solve():
    """qg_plan_probe (the plan arithmetic the handle uses, reachable without a GPU): per real column of the packed
    spectral layout the decay r of the y-recurrences equals the model's root (tests/ysolve_model.py) for that
    column's (field, wavenumber); kappa = -r dx^2 / M; the singular Poisson k = 0 column is marked by r = 0.  The
    work list of the y-slab edge correction is exactly the set of (32-column tile, 32-row segment) pairs closer
    than 41.6 / -ln r rows to a slab edge for the slowest-decaying column of the tile."""
    import ysolve_model as ym
    _ensure_built()
    lib = qgb200.load()
    for M, P, rows in ((64, 64, 32), (256, 512, 128), (2048, 4096, 512)):
        m = qgb200.BaroclinicModel(1000., 2000., 2e-11, 4e6, 4e6 * P / M, 300., 86400., 0.1, M, P, 4e6 / M, 100., 1e-7, 4e4, 1e-6)
        p = qgb200.make_params(m)
        ncol = 2 * M
        r = (ctypes.c_double * ncol)()
        kap = (ctypes.c_double * ncol)()
        n = ctypes.c_int(0)
        assert lib.qg_plan_probe(ctypes.byref(p), rows, r, kap, None, 0, ctypes.byref(n)) == 0
        work = (ctypes.c_int32 * (2 * n.value))()
        assert lib.qg_plan_probe(ctypes.byref(p), rows, None, None, work, n.value, ctypes.byref(n)) == 0
        r, kap = np.array(r[:]), np.array(kap[:])
        # the layout: slot s = col // 2; slot 0 and M/2 hold (field 0, field 1) of k = 0 / M/2, slot s < M/2 field 0 with
        # k = s, slot s > M/2 field 1 with k = M - s
        want = np.zeros(ncol)
        for col in range(ncol):
            s, part = col >> 1, col & 1
            if s == 0:
                field, k = part, 0
            elif s == M // 2:
                field, k = part, M // 2
            elif 2 * s < M:
                field, k = 0, s
            else:
                field, k = 1, M - s
            e = 4.0 * np.sin(np.pi * k / M) ** 2 - (p.alpha * m.dx * m.dx if field == 1 else 0.0)
            want[col] = ym.root(e) if e > 0 else 0.0
        assert r[0] == 0.0 and np.all(r[1:] > 0.0) and np.all(r < 1.0)
        assert np.allclose(r, want, rtol=2e-15, atol=0.0)
        assert np.allclose(kap, -r * m.dx * m.dx / M, rtol=2e-15, atol=0.0)
        pairs = {(work[2 * i], work[2 * i + 1]) for i in range(n.value)}
        expect = set()
        with np.errstate(divide="ignore"):
            ncut_col = np.where(r > 0, np.minimum(rows, np.floor(-41.6 / np.log(np.where(r > 0, r, 0.5))) + 1), 0)
        for t in range(ncol // 32):
            ncut = int(ncut_col[32 * t:32 * t + 32].max())
            for sg in range(rows // 32):
                if 32 * sg < ncut or rows - (32 * sg + 31) <= ncut:
                    expect.add((t, sg))
        assert pairs == expect and len(pairs) == n.value
        # every tile is touched at both edges; only the longest waves everywhere
        assert all((t, 0) in pairs and (t, rows // 32 - 1) in pairs for t in range(ncol // 32))
        if rows >= 128 and M >= 256:
            assert n.value < (ncol // 32) * (rows // 32)
    assert lib.qg_plan_probe(None, 32, None, None, None, 0, None) == -1
    assert lib.qg_plan_probe(ctypes.byref(p), 48, None, None, None, 0, ctypes.byref(n)) == -1
