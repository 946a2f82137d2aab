"""Host-side helpers for the y-slab decomposition of one run over several GPUs (new: the
reference is a single process).  Rank r of G owns global rows [r*P/G, (r+1)*P/G) with every x;
its host arrays are (M+2, P/G+2, 2, 3) whose ghost rows hold the neighbouring ranks' rows
(periodic ring), exactly what an MPI-style Julia caller would keep."""
from __future__ import annotations

import numpy as np


def row_range(P, rank, nranks):
    if P % nranks:
        raise ValueError("P must be divisible by the number of ranks")
    pl = P // nranks
    return rank * pl, (rank + 1) * pl


def local_model(model, nranks, cls=None):
    """The model a rank hands to qg_create: same parameters, P and Ly of its slab."""
    j0, j1 = row_range(model.P, 0, nranks)
    cls = cls or type(model)
    args = [getattr(model, k) for k in ("H_1", "H_2", "beta", "Lx", "Ly", "dt", "T", "U", "M", "P", "dx", "visc",
                                         "r", "R_d", "initial_kick")]
    args[9] = j1 - j0
    return cls(*args)


def take_slab(a, rank, nranks):
    """Slice a global (M+2, P+2, ...) array (ghost ring included) into rank's local array with
    ghost rows: global ghost-inclusive rows [j0, j1+2)."""
    P = a.shape[1] - 2
    j0, j1 = row_range(P, rank, nranks)
    return np.asfortranarray(a[:, j0:j1 + 2, ...])


def put_slab(a_global, a_local, rank, nranks):
    """Write a rank's interior rows back into the global array (ghost ring not touched)."""
    P = a_global.shape[1] - 2
    j0, j1 = row_range(P, rank, nranks)
    a_global[:, j0 + 1:j1 + 1, ...] = a_local[:, 1:-1, ...]
    return a_global


def refresh_global_ghosts(a):
    """Periodic ghost ring of a reassembled global array (all layers / levels)."""
    a[1:-1, 0, ...] = a[1:-1, -2, ...]
    a[1:-1, -1, ...] = a[1:-1, 1, ...]
    a[0, :, ...] = a[-2, :, ...]
    a[-1, :, ...] = a[1, :, ...]
    return a
