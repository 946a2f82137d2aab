"""Developer tool (torchrun, one rank per GPU): y-slab run compared with the oracle after EVERY step, per rank,
with the location of the largest error - to find where and when a slab run departs.
    torchrun ... scripts/slab_steps_debug.py M P steps [sync]"""
import datetime
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import qgb200
    M, P, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    sync = len(sys.argv) > 4
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    o, oc = bench.oracle_modules()
    sess, glob, a, peer = bench.slab_session(torch, dist, qgb200, stream, local, rank, world, M, P)
    ref = torch.empty((steps, 2, 2, P + 2, M + 2), dtype=torch.float64, device="cuda")
    if rank == 0:
        with qgb200.Session(glob, members=1, device=local, stream=stream.cuda_stream) as s1:
            s1.init_state(1)
            zeta, psi = s1.new_state_array(), s1.new_state_array()
            s1.download(zeta=zeta, psi=psi)
        mo = o.make_model(*[a[k] for k in bench.MODEL_KEYS])
        f = np.zeros_like(zeta)
        for t in range(steps):
            oc.run_steps(mo, zeta, psi, f, 1 + t, 1, bench.host_threads(oc))
            ref[t, 0].copy_(torch.from_numpy(np.ascontiguousarray(zeta[:, :, :, 0].T)))
            ref[t, 1].copy_(torch.from_numpy(np.ascontiguousarray(psi[:, :, :, 0].T)))
    dist.broadcast(ref, src=0)
    refh = ref.cpu().numpy()
    sess.init_state(1)
    if sync:
        sess.sync(); dist.barrier()
    pl_rows = P // world
    j0 = rank * pl_rows
    zl, pl = sess.new_state_array(), sess.new_state_array()
    for t in range(steps):
        sess.step(1 + t, 1)
        sess.download(zeta=zl, psi=pl)
        out = {"step": t + 1, "rank": rank}
        for k, (name, mine) in enumerate((("q", zl), ("psi", pl))):
            want = refh[t, k][:, j0:j0 + pl_rows + 2, :]
            got = np.ascontiguousarray(mine[:, :, :, 0].T)
            d = np.abs(got - want)
            l, jj, ii = np.unravel_index(int(np.argmax(d)), d.shape)
            rowmax = d.max(axis=(0, 2)) / np.abs(want).max()
            bad = np.nonzero(rowmax > 1e-10)[0]
            out[name] = {"err": float(d.max() / np.abs(want).max()), "at": [int(l), int(jj) - 1, int(ii) - 1],
                         "bad_rows": [int(x) - 1 for x in bad[:6]] + (["..."] if len(bad) > 6 else []), "n_bad_rows": int(len(bad))}
        allo = [None] * world
        dist.all_gather_object(allo, out)
        if rank == 0:
            for x in allo:
                print(json.dumps(x), flush=True)
    dist.barrier()
    sess.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
