#!/bin/bash
# developer script (run under gpurun --gpus N): y-slab parity over N ranks, strong scaling of
# 16384x8192 (peer-memory exchange and, with a second argument, the NCCL fallback), and the
# member-sharded ensembles.
N=${1:-4}; EXTRA=${2:-}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tests/dist_slab_check.py 1024 2048 10 2>&1 | grep -E "SLAB_CHECK|slab check|rror" | tail -3
slab() { timeout 400 $TR --master-port 29551 bench.py --gpus $N --mode slab --grid 16384 8192 --steps 50 --warmup 5 2>&1 | grep '^{' | tee gpurun_out/slab_16384x8192_n${N}$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['per_rank_kernels_us']
print(round(d['ms_per_step'],4), '%.3e'%d['value'], {n:v['us'] for n,v in k.items()})"; }
echo "== slab peer"; slab ""
if [ -n "$EXTRA" ]; then
  echo "== slab nccl"; QG_DIST_NCCL=1 slab _nccl
  echo "== config 5: 64 members of 512x512, 8 per GPU"
  timeout 300 $TR --master-port 29552 bench.py --gpus $N --grid 512 512 --members 8 --steps 300 --warmup 6 2>&1 | grep '^{' | tee gpurun_out/ensemble512_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
  echo "== 4096^2 per GPU (weak)"
  timeout 300 $TR --master-port 29553 bench.py --gpus $N --steps 100 --warmup 5 2>&1 | grep '^{' | tee gpurun_out/ensemble4096_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
fi
