"""y-slab parity check, run as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_slab_check.py [M P steps]

Every rank builds the same global seeded initial condition, takes its slab of rows, steps it
through the C ABI in y-slab mode (exchanges over NVLink peer memory; QG_DIST_NCCL=1: NCCL halo
exchange + carry all-gather), downloads, and compares
with the CPU oracle's solution of the GLOBAL problem restricted to the slab (rank 0 prints)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qg_oracle as o  # noqa: E402
import qg_oracle_c as oc  # noqa: E402
import qgb200  # noqa: E402
from qgb200 import slab  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # host-side plumbing only (the id broadcast); NCCL lives in libqgb200
    M, P, steps = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 256, 10)))
    mo = o.standard_model(M, P, dt=300.0 if M > 1024 else 3600.0)
    zeta, psi = o.initialise_model(mo, seed=1)
    f = np.zeros_like(zeta)
    ids = [qgb200.Session.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ml = slab.local_model(mo, world, cls=qgb200.BaroclinicModel)
    mo_g = slab.local_model(mo, 1, cls=qgb200.BaroclinicModel)   # the global grid as a product model
    ic_ok = True
    zl, pl, fl = (slab.take_slab(a, rank, world) for a in (zeta, psi, f))
    def gather_blobs(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    with qgb200.Session(ml, device=local) as s:
        s.dist_init(rank, world, ids[0])
        if not os.environ.get("QG_DIST_NCCL"):
            s.dist_peer_init(gather_blobs)   # per-step exchanges over NVLink peer memory instead of NCCL
        if M * P <= 1 << 22:
            # the device initial condition of a slab = the slab of the single-GPU initial condition, bit for bit
            s.init_state(11)
            za, pa = s.new_state_array(), s.new_state_array()
            s.download(zeta=za, psi=pa)
            with qgb200.Session(mo_g, device=local) as sg:
                sg.init_state(11)
                zg, pg = sg.new_state_array(), sg.new_state_array()
                sg.download(zeta=zg, psi=pg)
            ic_ok = np.array_equal(za, slab.take_slab(zg, rank, world)) and np.array_equal(pa, slab.take_slab(pg, rank, world))
        s.upload(zl, pl, fl)
        s.step(1, steps)
        s.download(zl, pl, fl)
        E, Z = s.diagnostics()
        if os.environ.get("QG_VERBOSE") and rank == 0:
            print(f"rank 0: {s.launch_count()} kernel launches", flush=True)
        dist.barrier()   # nobody frees memory a peer still has mapped
    errs = None
    if rank == 0 or True:
        if M * P <= 1 << 18:
            o.run_steps(mo, zeta, psi, f, o.make_factors(mo, "spectral"), 1, steps)
        else:   # the C restatement (pinned to the NumPy oracle in tests/test_oracle_c.py)
            oc.run_steps(mo, zeta, psi, f, 1, steps, max(1, (os.cpu_count() or 2) // world))
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        ref = [slab.take_slab(a, rank, world) for a in (zeta, psi, f)]
        Eo, Zo = o.diagnostics(mo, zeta, psi)
        errs = [rel(zl, ref[0]), rel(pl, ref[1]), rel(fl, ref[2]), abs(E - Eo) / Eo, abs(Z - Zo) / Zo,
                0.0 if ic_ok else 1.0]
    allerrs = [None] * world
    dist.all_gather_object(allerrs, errs)
    if rank == 0:
        worst = np.max(np.array(allerrs), axis=0)
        print(f"slab check {M}x{P} over {world} ranks, {steps} steps: q {worst[0]:.2e} psi {worst[1]:.2e} "
              f"f {worst[2]:.2e} E {worst[3]:.2e} Z {worst[4]:.2e}")
        ok = (worst[0] < 1e-10 and worst[1] < 1e-10 and worst[2] < 1e-10 and worst[3] < 1e-8 and worst[4] < 1e-8
              and worst[5] == 0.0)
        if worst[5] != 0.0:
            print("slab-mode qg_init_state differs from the slab of the single-GPU initial condition")
        print("SLAB_CHECK_OK" if ok else "SLAB_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
