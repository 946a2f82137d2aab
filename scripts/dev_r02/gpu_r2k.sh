#!/bin/bash
TAG=${1:-r02k}
mkdir -p gpurun_out
{
echo "== default x4"; for i in 1 2 3 4; do python scripts/ab_run.py 4096 4096 40; done
echo "== pf=6 x3"; for i in 1 2 3; do QG_RING_PF=6 python scripts/ab_run.py 4096 4096 40; done
echo "== ring off x3"; for i in 1 2 3; do QG_FFT_RING=0 python scripts/ab_run.py 4096 4096 40; done
} > gpurun_out/det_$TAG.log 2>&1
cat gpurun_out/det_$TAG.log
