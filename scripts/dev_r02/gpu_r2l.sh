#!/bin/bash
mkdir -p gpurun_out
{
python scripts/determinism.py 4096 1024 30 60
QG_FFT_RING=0 python scripts/determinism.py 4096 1024 30 60
QG_NO_GRAPH=1 python scripts/determinism.py 4096 1024 30 60
python scripts/determinism.py 4096 4096 43 12
QG_FFT_RING=0 python scripts/determinism.py 4096 4096 43 12
python scripts/determinism.py 1024 1024 30 100
} > gpurun_out/det_r02l.log 2>&1
cat gpurun_out/det_r02l.log
