#!/bin/bash
# developer script (run under gpurun): GPU tests + one bench line summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print(round(d['ms_per_step'],4), {n:(v['us'],v['frac']) for n,v in k.items()}, d['roofline']['small_kernels_us'], 'e2e', d['e2e']['value'])"; }
run QG_X=1
