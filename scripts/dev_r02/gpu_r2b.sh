#!/bin/bash
# developer script (1 GPU): ring-buffered FFT kernels vs the round-1 kernels, bit-identity + timing; new tests
TAG=${1:-r02b}
mkdir -p gpurun_out
{
for ring in 0 1; do
  QG_FFT_RING=$ring python scripts/ab_run.py 4096 4096 30
  QG_FFT_RING=$ring python scripts/ab_run.py 4096 300 10
done
( time timeout 900 python -m pytest tests -m gpu -x -q -k "4096 or monitor or reupload or 1000_steps_1024 or single_use or snapshots or initial_condition" ) 2>&1 | tail -6
for cs in 1 2; do
  QG_COPY_STREAMS=$cs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config4 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('copy_streams=$cs value %.4e ms %.4f graph %.4f e2e %.4e (%.2f ms)'%(d['value'],d['ms_per_step'],d['ms_per_step_graph_replay'],d['e2e']['value'],d['e2e']['ms_total']), {k:(v['us'],v['frac']) for k,v in d['roofline']['kernels'].items()})"
done
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
