"""Developer script (run under `compute-sanitizer --tool memcheck` on a GPU box): a few small steps
through every kernel family - persistent y-solve (ragged and full), radix-16 / radix-8 / long / DFT
row transforms, batched members, snapshots, device IC."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
import numpy as np
import qgb200

def model(M, P):
    dx = 4.0e6 / M
    return qgb200.BaroclinicModel(1e3, 2e3, 2e-11, 4.0e6, dx * P, 3600.0, 86400.0, 0.1, M, P, dx, 100.0, 1e-7, 4e4, 1e-6)

for M, P, nm in ((64, 64, 1), (256, 96, 2), (40, 24, 1), (128, 1056, 1), (16384, 32, 1), (64, 4160, 1)):
    m = model(M, P)
    with qgb200.Session(m, members=nm) as s:
        s.init_state(3)
        s.step(1, 7)
        z1 = np.zeros((M + 2, P + 2, 2) + ((nm,) if nm > 1 else ()), order="F")
        p1 = np.zeros_like(z1)
        s.snapshot_begin(z1, p1)
        s.step(8, 2)
        s.snapshot_end()
        E, Z = s.diagnostics()
        assert np.all(np.isfinite(E)) and np.all(np.isfinite(Z)) and np.isfinite(z1).all() and p1.any()
    print("ok", M, P, nm, flush=True)
