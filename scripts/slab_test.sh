#!/bin/bash
# developer script (run under gpurun --gpus 2): y-slab parity at several sizes + the slab bench
set -o pipefail
for sz in "256 256" "512 1024" "2048 4096" "16384 128"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tests/dist_slab_check.py $sz 10 2>&1 | grep -E "SLAB_CHECK|rel|Error|error" | tail -3
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus 2 --mode slab --grid 16384 8192 --steps 40 --warmup 5 2>&1 | grep '^{' | tee gpurun_out/slab_16384x8192_n2_new.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['per_rank_kernels_us']
print(round(d['ms_per_step'],4), '%.3e'%d['value'], {n:v['us'] for n,v in k.items()})"
