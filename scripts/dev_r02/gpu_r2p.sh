#!/bin/bash
# developer script (gpurun --gpus N): the contract line at N GPUs (top level: one 4096^2 run per GPU; `slab`: config 4
# in N y-slabs with parity) and, with a second argument, the full-size slab parity against the oracle
N=${1:-8}; FULL=${2:-}
TAG=r02p
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
grep '^{' gpurun_out/bench_n${N}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d.get('slab',{})
print('N=$N value %.4e e2e %.4e ms %.4f'%(d['value'], d['e2e']['value'], d['ms_per_step']))
print('slab', {k:s.get(k) for k in ('ms_per_step','value','speedup_vs_single_gpu','per_rank_kernels_us_max','barriers_and_gaps_us_rank0','error')})
print('parity', s.get('parity')); print('full', s.get('full_size_check')); print('single', (s.get('single_gpu') or {}).get('ms_per_step'))"
tail -3 gpurun_out/bench_n${N}_$TAG.err
if [ -n "$FULL" ]; then
  ( time timeout 900 $TR --master-port 29563 scripts/parity_large.py slab 16384 8192 10 ) > gpurun_out/parity_slab_n${N}_$TAG.log 2>&1
  grep -E "^\{|SLAB_CHECK" gpurun_out/parity_slab_n${N}_$TAG.log | cut -c1-600
  timeout 300 $TR --master-port 29533 tests/dist_slab_check.py 1024 2048 10 2>&1 | grep -E "SLAB_CHECK|slab check" | tee gpurun_out/slab_check_n${N}_$TAG.log
fi
