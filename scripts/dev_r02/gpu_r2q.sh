#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
check() { timeout 300 $TR --master-port 29533 tests/dist_slab_check.py "$@" 2>&1 | grep -E "SLAB_CHECK|slab check|Error|error|differs" | tail -3; }
{
echo "== single pass, $N ranks"; check 2048 $((512*N)) 10; check 2048 $((256*N)) 10; check 1024 $((512*N)) 10; check 4096 $((512*N)) 10
echo "== two pass"; QG_K3_TWOPASS=1 check 2048 $((512*N)) 10
} > gpurun_out/slab_r02q_n$N.log 2>&1
cat gpurun_out/slab_r02q_n$N.log
