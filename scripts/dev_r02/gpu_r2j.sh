#!/bin/bash
TAG=${1:-r02j}
mkdir -p gpurun_out
{
for pf in 0 3 6; do
  QG_RING_PF=$pf python scripts/ab_run.py 4096 4096 40
done
QG_RING_PF=3 python scripts/ab_run.py 4096 300 10
timeout 600 python -m pytest tests -m gpu -x -q -k "4096 or pair" 2>&1 | tail -3
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
