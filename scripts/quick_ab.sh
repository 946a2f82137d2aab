#!/bin/bash
# developer script (run under gpurun): bench line summaries for the smaller BASELINE configs
run() { echo "== $*"; python bench.py "$@" --no-cpu-baseline 2> gpurun_out/ab.err | tee -a gpurun_out/small_grids.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print(round(d['ms_per_step']*1e3,1),'us/step', '%.3e'%d['value'], {n:v['us'] for n,v in k.items()}, d['roofline']['small_kernels_us'], 'e2e %.3e'%d['e2e']['value'])"; tail -2 gpurun_out/ab.err; }
rm -f gpurun_out/small_grids.jsonl
run --grid 1024 1024 --steps 300 --warmup 6
run --grid 512 512 --members 8 --steps 300 --warmup 6
run --grid 128 128 --steps 300 --warmup 6
