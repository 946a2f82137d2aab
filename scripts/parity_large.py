"""Parity at the sizes the correctness gates are stated on (BASELINE.json north_star):
psi and q <= 1e-10 relative after 10 steps, energy and enstrophy <= 1e-8 relative after 1000 steps,
against the CPU oracle (oracle/qg_oracle.c, the C restatement pinned to the NumPy one in
tests/test_oracle_c.py) from the SAME initial condition (drawn on the device, downloaded for the oracle).

    python scripts/parity_large.py single 4096 4096 1000          # config 3, 1000 steps (~7 min of oracle)
    python scripts/parity_large.py single 16384 8192 10           # config 4 grid on one GPU
    torchrun --nproc-per-node N scripts/parity_large.py slab 16384 8192 10   # config 4 in N y-slabs

Each run prints one JSON line (kept under profiles/).  Test infrastructure, not product code."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (model parameters, oracle loader and the slab parity leg are shared with bench.py)


def single(M, P, steps):
    import numpy as np
    import qgb200
    o, oc = bench.oracle_modules()
    model, a = bench.make_model(qgb200, M, P)
    mo = o.make_model(*[a[k] for k in bench.MODEL_KEYS])
    t0 = time.perf_counter()
    with qgb200.Session(model) as s:
        s.init_state(1)
        zeta, psi = s.new_state_array(), s.new_state_array()
        s.download(zeta=zeta, psi=psi)
        s.step(1, steps)
        z, p = s.new_state_array(), s.new_state_array()
        s.download(zeta=z, psi=p)
        E, Z = s.diagnostics()
    t_gpu = time.perf_counter() - t0
    f = np.zeros_like(zeta)
    t0 = time.perf_counter()
    oc.run_steps(mo, zeta, psi, f, 1, steps, bench.host_threads(oc))
    t_cpu = time.perf_counter() - t0
    Eo, Zo = o.diagnostics(mo, zeta, psi)
    rel = lambda x, y: float(np.abs(x - y).max() / np.abs(y).max())
    out = {"case": f"single GPU {M}x{P}, {steps} steps, dt {a['dt']} s",
           "q": max(rel(z[:, :, l, 0], zeta[:, :, l, 0]) for l in range(2)),
           "psi": max(rel(p[:, :, l, 0], psi[:, :, l, 0]) for l in range(2)),
           "q_older_levels": max(rel(z[:, :, l, k], zeta[:, :, l, k]) for l in range(2) for k in (1, 2)),
           "psi_older_levels": max(rel(p[:, :, l, k], psi[:, :, l, k]) for l in range(2) for k in (1, 2)),
           "E": abs(E - Eo) / abs(Eo), "Z": abs(Z - Zo) / abs(Zo), "E_gpu": E, "E_oracle": float(Eo),
           "gpu_wall_s": round(t_gpu, 2), "oracle_wall_s": round(t_cpu, 2)}
    field_tol = 1e-10 if steps <= 10 else 1e-8   # the field gate is stated for 10 steps; E/Z for 1000
    out["ok"] = bool(out["q"] <= field_tol and out["psi"] <= field_tol and out["E"] <= 1e-8 and out["Z"] <= 1e-8)
    out["tolerance"] = f"q, psi <= {field_tol:g}; E, Z <= 1e-8"
    print(json.dumps(out), flush=True)
    print("PARITY_LARGE_OK" if out["ok"] else "PARITY_LARGE_FAILED", flush=True)


def slab(M, P, steps):
    import datetime
    import numpy as np
    import torch
    import torch.distributed as dist
    import qgb200
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=1200))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = bench.slab_parity_leg(torch, dist, qgb200, np, stream, local, rank, world, steps=steps, grid=(M, P))
    if rank == 0:
        out["case"] = f"{world} y-slabs, {M}x{P}, {steps} steps"
        print(json.dumps(out), flush=True)
        print("SLAB_CHECK_OK" if out["ok"] else "SLAB_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    mode, M, P, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    (single if mode == "single" else slab)(M, P, steps)
