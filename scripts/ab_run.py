"""Developer tool: run `steps` steps of an M x P grid from the device initial condition and print a
SHA-256 of the downloaded (zeta, psi) plus per-kernel times - run it twice with different
environment switches (e.g. QG_FFT_RING=0 / 1) to check that two kernel variants are bit-identical."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qgb200  # noqa: E402

M, P, steps = (int(x) for x in sys.argv[1:4])
model, _ = bench.make_model(qgb200, M, P)
with qgb200.Session(model) as s:
    s.init_state(3)
    s.step(1, 3)
    s.set_profiling(True)
    s.step(4, steps)
    kt = s.kernel_times()
    s.set_profiling(False)
    z, p = s.new_state_array(), s.new_state_array()
    s.download(zeta=z, psi=p)
h = hashlib.sha256(z.tobytes() + p.tobytes()).hexdigest()[:16]
print(f"AB {M}x{P} {steps} steps sha {h} " + " ".join(f"{k}={ms / n * 1e3:.1f}us" for k, (ms, n) in kt.items() if n))
