#!/bin/bash
# developer script: y-slab strong scaling of one run (run under gpurun --gpus N)
N=${1:-2}; M=${2:-16384}; P=${3:-8192}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus $N --mode slab --grid $M $P --steps 50 --warmup 5 2>&1 | grep '^{' | tee gpurun_out/slab_${M}x${P}_n${N}.json
