#!/bin/bash
mkdir -p gpurun_out
{
QG_FFT_RING=0 python scripts/determinism.py 4096 4096 43 36
QG_FFT_RING=0 QG_K3_V1=1 python scripts/determinism.py 4096 4096 43 36
QG_FFT_RING=0 QG_FFT_PF=0 python scripts/determinism.py 4096 4096 43 36
python scripts/determinism.py 4096 4096 43 12
} > gpurun_out/det_r02n.log 2>&1
cat gpurun_out/det_r02n.log
