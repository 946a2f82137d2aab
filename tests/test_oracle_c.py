"""The C restatement (oracle/qg_oracle.c, used as CPU baseline) against the NumPy oracle."""
import numpy as np
import pytest

import qg_oracle as o
import qg_oracle_c as oc


@pytest.mark.parametrize("M,P,backend", [(8, 8, "direct"), (16, 8, "direct"), (24, 40, "direct"), (9, 7, "direct"),
                                         (64, 64, "spectral"), (128, 96, "spectral")])
def test_c_port_matches_numpy_oracle(M, P, backend):
    m = o.standard_model(M, P)
    z, p = o.initialise_model(m, seed=1)
    f = np.zeros_like(z)
    zc, pc, fc = z.copy(order="F"), p.copy(order="F"), f.copy(order="F")
    o.run_steps(m, z, p, f, o.make_factors(m, backend), 1, 10)
    oc.run_steps(m, zc, pc, fc, 1, 10, nthreads=2)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(pc, p) < 1e-11 and rel(zc, z) < 1e-12 and rel(fc, f) < 1e-11
