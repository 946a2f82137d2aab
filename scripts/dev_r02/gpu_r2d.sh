#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q -k "16384-3" 2>&1 | grep -E "assert|Error|error|rel\(" | head -20
QG_FFT_PAIR=0 timeout 120 python - <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'julia-ocean-modelling_b200/python')
import numpy as np, qg_oracle as o, qgb200
import test_gpu_parity as t
for (M,P) in ((16384,3),(16384,32)):
    mo, mg = t.models(M, P)
    zeta, psi = o.initialise_model(mo, seed=1); f = np.zeros_like(zeta)
    z, p, ff, _, _ = t.gpu_run(mg, zeta, psi, f, 1, 10)
    o.run_steps(mo, zeta, psi, f, o.make_factors(mo, "spectral"), 1, 10)
    print("old kernels", M, P, t.rel(p, psi), t.rel(z, zeta))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rfft_pair' -s 8 -c 2 -o gpurun_out/prof_pair_$TAG -f python scripts/ab_run.py 16384 2048 4 2>&1 | tail -3
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
