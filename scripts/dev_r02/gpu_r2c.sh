#!/bin/bash
# developer script (1 GPU): cluster-pair long-row transforms vs the single-CTA kernels
TAG=${1:-r02c}
mkdir -p gpurun_out
{
( time timeout 600 python -m pytest tests -m gpu -x -q -k "16384" ) 2>&1 | tail -8
for pair in 0 1; do
  QG_FFT_PAIR=$pair timeout 300 python scripts/ab_run.py 16384 2048 20
done
QG_FFT_PAIR=1 QG_FFT_PF=0 timeout 300 python scripts/ab_run.py 16384 2048 20
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
