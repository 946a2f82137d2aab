"""Developer tool: repeat the same run R times in one process and count distinct result hashes (a kernel
race shows up as more than one).  python scripts/determinism.py M P steps R"""
import collections
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qgb200  # noqa: E402

M, P, steps, R = (int(x) for x in sys.argv[1:5])
model, _ = bench.make_model(qgb200, M, P)
seen = collections.Counter()
with qgb200.Session(model) as s:
    z, p = s.new_state_array(), s.new_state_array()
    for r in range(R):
        s.init_state(3)
        s.step(1, steps)
        s.download(zeta=z, psi=p)
        seen[hashlib.sha256(z.tobytes() + p.tobytes()).hexdigest()[:12]] += 1
print(f"DET {M}x{P} {steps} steps x{R}: {dict(seen)}", {k: os.environ.get(k) for k in ("QG_FFT_RING", "QG_NO_GRAPH", "QG_RING_PF")})
