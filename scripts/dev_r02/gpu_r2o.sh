#!/bin/bash
mkdir -p gpurun_out
{
python scripts/determinism.py 4096 4096 43 60
QG_FFT_RING=2 python scripts/determinism.py 4096 4096 43 24
python scripts/ab_run.py 4096 4096 40
python scripts/determinism.py 16384 2048 13 12
} > gpurun_out/det_r02o.log 2>&1
cat gpurun_out/det_r02o.log
