#!/bin/bash
# developer script (run under gpurun): A/B of the y-solve variants + phase trace
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print(round(d['ms_per_step'],4), {n:(v['us'],v['frac']) for n,v in k.items()}, d['roofline']['small_kernels_us'])"; grep qgb200 gpurun_out/ab.err | head -2; }
run QG_K3_V1=1
run QG_VERBOSE=1
run QG_K3_NCL=32
run QG_K3_NCL=16
QGB200_LIB=$PWD/julia-ocean-modelling_b200/lib_trace/libqgb200.so python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | grep K3TRACE
