#!/bin/bash
TAG=${1:-r02h}
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
check() { timeout 300 $TR --master-port 29533 tests/dist_slab_check.py "$@" 2>&1 | grep -E "SLAB_CHECK|slab check|Error|error|differs" | tail -4; }
{
echo "== single"; check 16384 128 10; check 16384 512 10; check 8192 256 10
echo "== two pass"; QG_K3_TWOPASS=1 check 16384 128 10; QG_K3_TWOPASS=1 check 16384 512 10; QG_K3_TWOPASS=1 check 8192 256 10
} > gpurun_out/acc_$TAG.log 2>&1
cat gpurun_out/acc_$TAG.log
