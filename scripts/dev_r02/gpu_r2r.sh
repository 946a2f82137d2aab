#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
for g in "2048 $((512*N))" "2048 $((1024*N))" "1024 $((512*N))"; do
  timeout 300 $TR --master-port 29563 scripts/parity_large.py slab $g 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-1500
done
QG_K3_TWOPASS=1 timeout 300 $TR --master-port 29563 scripts/parity_large.py slab 2048 $((512*N)) 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-1500
} > gpurun_out/slab_r02r_n$N.log 2>&1
cat gpurun_out/slab_r02r_n$N.log
