#!/bin/bash
# developer script (run under gpurun): bench line summaries under different env settings
run() { echo "== $*"; env "$@" python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print(round(d['ms_per_step'],4), {n:(v['us'],v['frac']) for n,v in k.items()}, d['roofline']['small_kernels_us'])"; tail -2 gpurun_out/ab.err; }
run QG_FFT_MINB=2
run QG_FFT_MINB=3
