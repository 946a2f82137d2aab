#!/bin/bash
# developer script: K1 tile-height sweep (run under gpurun)
for ty in 8 12 16 24; do
  echo "QG_K1_TY=$ty"
  QG_K1_TY=$ty python bench.py --steps 60 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernels']['k1_zeta_step'])"
done
