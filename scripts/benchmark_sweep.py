"""Counterpart of the reference's benchmark scripts (src/benchmarking/benchmarking.jl:28-44 and
src/benchmarking/julia_bench_parts.jl:30-52): for each M in the sweep, time the whole
run_model_no_output call, one evolve_psi, one evolve_zeta (timestep 1, as the reference does) and
the plan construction (the stand-in for the two Cholesky factorisations), and write the same CSV
columns so the results can be plotted next to the reference's.

    python scripts/benchmark_sweep.py [--parts] [--out file.csv]
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
    args = ap.parse_args()
    Ms = list(range(8, 129, 8)) if args.parts else [8, 16, 32, 64, 128]
    dt = (30.0 if args.parts else 60.0) * MINUTES
    rows = []
    for M in Ms:
        m = model(M, dt)
        r = (np.asfortranarray(np.random.default_rng(1).random((M + 2, M + 2))),
             np.asfortranarray(np.random.default_rng(2).random((M + 2, M + 2))))
        qgb200.run_model_no_output(m, rand_fields=r)   # warm-up (library load, plan caches)
        total = best(lambda: qgb200.run_model_no_output(m, rand_fields=r), args.samples)
        row = {"M": M, "total_time" if args.parts else "Time": total}
        if args.parts:
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
