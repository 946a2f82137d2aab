#!/bin/bash
# developer script (run under gpurun --gpus N): where does the y-slab step spend its non-kernel time?
N=${1:-2}; M=${2:-16384}; P=${3:-8192}
for mask in 0 7 1 2 4; do
  echo "== QG_DIST_SKIP=$mask"
  QG_DIST_SKIP=$mask python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
     bench.py --gpus $N --mode slab --grid $M $P --steps 40 --warmup 5 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['per_rank_kernels_us']
print(round(d['ms_per_step'],4), 'kernel sum', round(sum(v['us'] for v in k.values())/1e3,4), {n:v['us'] for n,v in k.items()})"
done
