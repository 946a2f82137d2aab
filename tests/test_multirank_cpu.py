"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab partitioning of the
reference-layout arrays and the rank-level algebra of the y-solve carry exchange, the same
(FF, RR, X, Y) all-gather + cyclic closure that qg_dist.cu / k3_rank_closure run over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ysolve_model as ym  # noqa: E402


def direct(d, g):
    P = len(g)
    A = np.zeros((P, P))
    for j in range(P):
        A[j, j] += d
        A[j, (j - 1) % P] += 1
        A[j, (j + 1) % P] += 1
    return np.linalg.solve(A, g)


@pytest.mark.parametrize("P,e", [(96, 0.3), (256, 6e-4), (100, 2.0), (33, 0.05)])
def test_chunked_factorised_solve_single_rank(P, e):
    g = np.random.default_rng(1).standard_normal(P)
    u = ym.solve_cyclic(g, e)
    ref = direct(-(2 + e), g)
    assert np.abs(u - ref).max() / np.abs(ref).max() < 1e-9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, e, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "julia-ocean-modelling_b200", "python"))
    from qgb200 import slab
    g = np.random.default_rng(7).standard_normal(P)      # same global rhs on every rank
    j0, j1 = slab.row_range(P, rank, world)
    r = ym.root(e)
    items, chunks = ym.slab_items(g[j0:j1], r)
    mine = ym.aggregate(items)                            # this rank's (FF, RR, X, Y)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                # the carry all-gather
    As, Be = ym.close_cyclic(gathered, 1.0 / (1.0 - r ** P))   # k3_rank_closure
    cAs, cBe = ym.close_open(items, As[rank], Be[rank])   # chunk level with the ends handed in
    u = ym.apply(g[j0:j1], r, chunks, cAs, cBe)
    # single-pass variant (k3_ysolve_pipe<3> + k3_rank_correct): solve with zero incoming carries, then add
    # the carry terms near the slab edges
    zAs, zBe = ym.close_open(items, 0.0, 0.0)
    u1 = ym.apply(g[j0:j1], r, chunks, zAs, zBe) + ym.rank_correction(j1 - j0, r, As[rank], Be[rank])
    assert np.abs(u1 - u).max() <= 1e-13 * np.abs(u).max(), np.abs(u1 - u).max() / np.abs(u).max()
    # slab partition of a reference-layout state array and its reassembly
    full = np.asfortranarray(np.random.default_rng(3).standard_normal((10, P + 2, 2, 3)))
    slab.refresh_global_ghosts(full)
    loc = slab.take_slab(full, rank, world)
    parts = [None] * world
    dist.all_gather_object(parts, (u, loc))
    if rank == 0:
        usol = np.concatenate([p[0] for p in parts])
        rebuilt = np.zeros_like(full)
        for rk, p in enumerate(parts):
            slab.put_slab(rebuilt, p[1], rk, world)
        slab.refresh_global_ghosts(rebuilt)
        halo_ok = all(np.array_equal(parts[rk][1][:, -1], parts[(rk + 1) % world][1][:, 1]) and
                      np.array_equal(parts[rk][1][:, 0], parts[(rk - 1) % world][1][:, -2]) for rk in range(world))
        q.put((usol, np.array_equal(rebuilt, full), halo_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,e", [(128, 0.2), (256, 6e-4), (512, 4.0)])
def test_two_rank_carry_exchange_matches_global_solve(P, e):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, e, q)) for r in range(2)]
    for p in procs:
        p.start()
    usol, rebuilt_ok, halo_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.random.default_rng(7).standard_normal(P)
    ref = direct(-(2 + e), g)
    assert np.abs(usol - ref).max() / np.abs(ref).max() < 1e-9
    assert rebuilt_ok and halo_ok
