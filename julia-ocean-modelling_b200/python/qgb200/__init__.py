"""qgb200 — B200-native Phillips two-layer QG time stepper (Python host twin of the Julia shim).

The package mirrors the reference's model API (``BaroclinicModel``, ``initialise_model``,
``evolve_zeta``, ``evolve_psi``, ``run_model_no_output`` ...) and forwards the hot path to
``libqgb200.so`` through ctypes.  Importing it does not need a GPU; creating a session does.
"""
from ._lib import LIB_PATH, QGError, load, qg_params  # noqa: F401
from .model import (DAY, KM, MINUTES, YEAR, BaroclinicModel, P_inv_matrix, P_matrix,  # noqa: F401
                    RectangularDomain, S1_plus, S2_minus, S_eig, Session, SpectralPlan, beta_1, beta_2,
                    close_sessions, evolve_psi, evolve_zeta, get_helmholtz_cholesky, get_poisson_cholesky,
                    initialise_model, make_params, ratio_term, run_model_no_output,
                    sp_solve_modified_helmholtz, sp_solve_poisson, update_doubly_periodic_bc)
from .runs import (MONITOR_COLUMNS, create_metadata, load_restart, load_run, log_model_params,  # noqa: F401
                   resume_model, run_model, save_restart, update_max, update_min)
from . import slab  # noqa: E402,F401
