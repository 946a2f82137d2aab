#!/bin/bash
# developer script (2 GPUs): single-pass y-slab solve (k3_ysolve_pipe<3> + k3_rank_correct) vs the two-pass flow
TAG=${1:-r02e}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
check() { timeout 300 $TR --master-port 29533 tests/dist_slab_check.py "$@" 2>&1 | grep -E "SLAB_CHECK|slab check|Error|error|differs" | tail -4; }
slab() { timeout 400 $TR --master-port 29551 bench.py --gpus $N --mode slab --grid 16384 8192 --steps 30 --warmup 5 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['slab']
print(round(d['ms_per_step'],4), '%.3e'%d['value'], s['per_rank_kernels_us_max'], 'gaps', s['barriers_and_gaps_us_rank0'], 'E', s['E'], 'Z', s['Z'])"; }
{
echo "== single pass"; check 256 256 10; check 512 1024 10; check 2048 4096 10; check 16384 128 10; QG_DIST_NCCL=1 check 512 1024 10
echo "== two pass";  QG_K3_TWOPASS=1 check 512 1024 10
echo "== bench single pass"; slab
echo "== bench two pass"; QG_K3_TWOPASS=1 slab
} > gpurun_out/slab_$TAG.log 2>&1
cat gpurun_out/slab_$TAG.log
