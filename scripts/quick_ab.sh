#!/bin/bash
# developer script (run under gpurun): long-row (M = 16384) transforms with / without L2 prefetch
run() { echo "== $*"; env "$@" python bench.py --grid 16384 1024 --steps 40 --warmup 5 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print(round(d['ms_per_step'],4), {n:(v['us'],v['frac']) for n,v in k.items()}, d['roofline']['small_kernels_us'])"; tail -2 gpurun_out/ab.err; }
run QG_FFT_PF=0
run QG_FFT_PF=1
