"""Developer script (GPU box): where the end-to-end time of bench.py's e2e leg goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, ROOT)
import numpy as np, torch
import qgb200, bench
M = P = 4096
a = bench.model_args(M, P)
model = qgb200.BaroclinicModel(*[a[k] for k in ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx", "visc", "r", "R_d", "initial_kick")])
zeta, psi = qgb200.initialise_model(model, seed=1)
n = zeta.size
pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
v = [p.numpy().reshape(zeta.shape, order="F") for p in pin]
v[0][...] = zeta; v[1][...] = psi; v[2][...] = 0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
s = qgb200.Session(model, members=1, device=0, stream=stream.cuda_stream)
def t(f, *args):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(*args); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
s.upload_raw(pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr()); s.step(1, 12)
for rep in range(2):
    up = t(s.upload_initial_raw, pin[0].data_ptr(), pin[1].data_ptr())
    st = t(s.step, 1, 200)
    dz = t(s.download_raw, pin[0].data_ptr(), 0, 0)
    dp = t(s.download_raw, 0, pin[1].data_ptr(), 0)
    print(f"upload_initial {up:.2f} ms ({2*n/3*8/up/1e6:.1f} GB/s)  200 steps {st:.2f} ms  download zeta {dz:.2f} ms ({n*8/dz/1e6:.1f} GB/s)  psi {dp:.2f} ms")
# raw copy speed for reference
d = torch.empty(n, dtype=torch.float64, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); pin[0].copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pin[0], non_blocking=True); torch.cuda.synchronize(); dt2 = time.perf_counter() - t0
    print(f"raw D2H {n*8/dt/1e9:.1f} GB/s  H2D {n*8/dt2/1e9:.1f} GB/s")
s.close()
