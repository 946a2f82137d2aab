#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
timeout 200 $TR --master-port 29563 scripts/parity_large.py slab 2048 4096 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-2500
QG_K3_TWOPASS=1 timeout 200 $TR --master-port 29564 scripts/parity_large.py slab 2048 4096 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-2500
QG_DIST_NCCL=1 timeout 200 $TR --master-port 29565 scripts/parity_large.py slab 2048 4096 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-2500
} > gpurun_out/slab_r02s_n8.log 2>&1
cat gpurun_out/slab_r02s_n8.log
