#!/bin/bash
mkdir -p gpurun_out
{
python scripts/determinism.py 4096 4096 43 36
QG_RING_PF=0 python scripts/determinism.py 4096 4096 43 36
QG_RING_DBG=1 python scripts/determinism.py 4096 4096 43 36
QG_RING_DBG=3 python scripts/determinism.py 4096 4096 43 36
} > gpurun_out/det_r02m.log 2>&1
cat gpurun_out/det_r02m.log
