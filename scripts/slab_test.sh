#!/bin/bash
# developer script (run under gpurun --gpus N): y-slab parity at several sizes + the slab bench,
# peer-memory exchange (default) and NCCL (QG_DIST_NCCL=1)
N=${1:-2}
check() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tests/dist_slab_check.py "$@" 2>&1 | grep -E "SLAB_CHECK|slab check|Error|error|rror" | tail -4; }
bench() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus $N --mode slab --grid 16384 8192 --steps 40 --warmup 5 2>&1 | grep '^{' | tee -a gpurun_out/slab_new_n$N.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['per_rank_kernels_us']
print(round(d['ms_per_step'],4), '%.3e'%d['value'], {n:v['us'] for n,v in k.items()})"; }
rm -f gpurun_out/slab_new_n$N.jsonl
check 256 256 10
check 512 1024 10
check 2048 4096 10
check 16384 128 10
QG_DIST_NCCL=1 check 256 256 10
echo "== peer"; bench
echo "== nccl"; QG_DIST_NCCL=1 bench
