"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/qg_oracle.py).

The reference is Julia and cannot run in this image, and its own tests hold no golden
trajectory, so these fixtures are produced by the oracle's *direct*
back-end (SuperLU on the reference's matrices) on seeded initial conditions.  They pin the
oracle against regressions and give the GPU tests a trajectory to compare with without
re-running the sparse factorisation.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import qg_oracle as o  # noqa: E402

CASES = [  # name, M, P, dt, steps, kick
    ("traj_8x8_s10", 8, 8, 3600.0, 10, 1e-6),
    ("traj_16x8_s10", 16, 8, 3600.0, 10, 1e-6),
    ("traj_24x40_s10", 24, 40, 3600.0, 10, 1e-6),
    ("traj_64x64_s10", 64, 64, 3600.0, 10, 1e-6),
]


def main():
    for name, M, P, dt, steps, kick in CASES:
        m = o.standard_model(M, P, dt=dt, initial_kick=kick)
        zeta, psi = o.initialise_model(m, seed=1)
        f = np.zeros_like(zeta)
        o.run_steps(m, zeta, psi, f, o.make_factors(m, "direct"), 1, steps)
        E, Z = o.diagnostics(m, zeta, psi)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), M=M, P=P, dt=dt, steps=steps, kick=kick, seed=1,
                            zeta=zeta, psi=psi, f_store=f, E=E, Z=Z)
        print(name, E, Z)
    # long runs: diagnostics only (1000 steps, the energy / enstrophy gate of BASELINE.json)
    rows = []
    for M, P, dt, kick in [(64, 64, 3600.0, 1e-6), (128, 128, 3600.0, 1e-6), (128, 64, 1800.0, 1e-2)]:
        m = o.standard_model(M, P, dt=dt, initial_kick=kick)
        zeta, psi = o.initialise_model(m, seed=1)
        f = np.zeros_like(zeta)
        o.run_steps(m, zeta, psi, f, o.make_factors(m, "direct"), 1, 1000)
        E, Z = o.diagnostics(m, zeta, psi)
        rows.append((M, P, dt, kick, 1000, E, Z))
        print(rows[-1])
    np.save(os.path.join(HERE, "diag_1000steps.npy"), np.array(rows))


if __name__ == "__main__":
    main()
