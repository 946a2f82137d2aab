"""Counterpart of the reference's benchmark scripts (src/benchmarking/benchmarking.jl:28-44 and
src/benchmarking/julia_bench_parts.jl:30-52): for each M in the sweep, time the whole
run_model_no_output call, one evolve_psi, one evolve_zeta (timestep 1, as the reference does) and
the plan construction (the stand-in for the two Cholesky factorisations), and write the same CSV
columns so the results can be plotted next to the reference's.

    python scripts/benchmark_sweep.py [--parts] [--out file.csv] [--reference-columns]

--reference-columns adds, side by side, BASELINE.md's B1 stand-in for the reference (Julia is not in the
image): the NumPy restatement with SciPy SuperLU on the reference's own matrices (oracle/qg_oracle.py,
single thread like the reference), timed the way the reference's harness times itself - the whole
run_model_no_output call (`ref_Time` / `ref_total_time`: initial condition + two factorisations + loop),
the loop alone (`ref_loop_time`), and in --parts mode one evolve_psi!, one evolve_zeta! and each
factorisation.  This is a CPU-baseline leg (the only use of oracle/ outside tests/ and bench.py's
cpu_baseline), never part of the product path.
"""
import argparse
import csv
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
import qgb200  # noqa: E402
from qgb200 import DAY, KM, MINUTES  # noqa: E402


def model(M, dt):
    Lx = 4000.0 * KM   # src/benchmarking/benchmarking.jl:6-21
    return qgb200.BaroclinicModel(1.0 * KM, 2.0 * KM, 2e-11, Lx, Lx, dt, 1.0 * DAY, 0.1, M, M, Lx / M, 100.0, 1e-7,
                                  40.0 * KM, 1e-6)


def best(fn, n):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", action="store_true", help="julia_bench_parts.jl: M = 8:8:128, dt = 30 min, per-part times")
    ap.add_argument("--out", default=None)
    ap.add_argument("--samples", type=int, default=20)
    ap.add_argument("--reference-columns", action="store_true")
    ap.add_argument("--reference-only", action="store_true",
                    help="only the B1 reference-algorithm columns (runs without a GPU)")
    args = ap.parse_args()
    o = None
    if args.reference_columns or args.reference_only:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import qg_oracle as o
    Ms = list(range(8, 129, 8)) if args.parts else [8, 16, 32, 64, 128]
    dt = (30.0 if args.parts else 60.0) * MINUTES
    rows = []
    for M in Ms:
        m = model(M, dt)
        r = (np.asfortranarray(np.random.default_rng(1).random((M + 2, M + 2))),
             np.asfortranarray(np.random.default_rng(2).random((M + 2, M + 2))))
        total = None
        row = {"M": M}
        if not args.reference_only:
            qgb200.run_model_no_output(m, rand_fields=r)   # warm-up (library load, plan caches)
            total = best(lambda: qgb200.run_model_no_output(m, rand_fields=r), args.samples)
            row["total_time" if args.parts else "Time"] = total
        if args.parts and not args.reference_only:
            zeta, psi = qgb200.initialise_model(m, rand_fields=r)
            f = np.zeros_like(zeta)
            with qgb200.Session(m) as s:
                s.upload(zeta, psi, f)

                def psi_call():
                    s.evolve_psi(); s.sync()

                def zeta_call():
                    s.evolve_zeta(1); s.sync()
                row["psi_time"] = best(psi_call, args.samples)
                row["zeta_time"] = best(zeta_call, args.samples)

            def plan():
                qgb200.Session(m).close()
            t = best(plan, 5)
            row["helmholtz_time"] = t   # one spectral plan replaces both factorisations
            row["poisson_times"] = 0.0
        if o is not None:   # B1: the reference algorithm (NumPy + SuperLU), one thread
            mo = o.make_model(m.H_1, m.H_2, m.beta, m.Lx, m.Ly, m.dt, m.T, m.U, m.M, m.P, m.dx, m.visc, m.r, m.R_d,
                              m.initial_kick)
            steps = int(np.floor(mo.T / mo.dt))
            nref = 3 if M <= 64 else 1
            row["ref_total_time" if args.parts else "ref_Time"] = best(
                lambda: o.run_model_no_output(mo, seed=1, backend="direct"), nref)
            zo, po = o.initialise_model(mo, seed=1)
            fo = np.zeros_like(zo)
            fac = o.make_factors(mo, "direct")
            t0 = time.perf_counter()
            o.run_steps(mo, zo, po, fo, fac, 1, steps)
            row["ref_loop_time"] = time.perf_counter() - t0
            row["ref_cell_steps_per_s"] = M * M * steps / row["ref_loop_time"]
            if total is not None:
                row["gpu_speedup_whole_run"] = row["ref_total_time" if args.parts else "ref_Time"] / total
            if args.parts:
                row["ref_psi_time"] = best(lambda: o.evolve_psi(mo, zo, po, *fac), 3)
                row["ref_zeta_time"] = best(lambda: o.evolve_zeta(mo, zo, po, 1, fo), 3)
                row["ref_helmholtz_time"] = best(lambda: o.get_helmholtz_cholesky(M, M, mo.dx, o.S_eig(mo)), 1)
                row["ref_poisson_times"] = best(lambda: o.get_poisson_cholesky(M, M, mo.dx), 1)
        rows.append(row)
        print(row, flush=True)
    out = args.out or ("qgb200_parts_benchmark.csv" if args.parts else "qgb200_benchmark_times.csv")
    with open(out, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0]))
        w.writeheader()
        w.writerows(rows)
    print("wrote", out)


if __name__ == "__main__":
    main()
