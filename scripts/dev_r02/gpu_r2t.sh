#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
echo "== no sync"; timeout 200 $TR --master-port 29563 scripts/slab_steps_debug.py 2048 4096 3 2>&1 | grep -E "^\{|rror" | cut -c1-400
echo "== sync after init"; timeout 200 $TR --master-port 29564 scripts/slab_steps_debug.py 2048 4096 2 sync 2>&1 | grep -E "^\{|rror" | cut -c1-400
} > gpurun_out/slab_r02t_n8.log 2>&1
cat gpurun_out/slab_r02t_n8.log
