#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
timeout 150 $TR --master-port 29563 scripts/parity_large.py slab 2048 4096 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-900
timeout 150 $TR --master-port 29564 scripts/parity_large.py slab 4096 4096 10 2>&1 | grep -E "^\{|SLAB_CHECK|rror" | cut -c1-500
} > gpurun_out/slab_r02u_n8.log 2>&1
cat gpurun_out/slab_r02u_n8.log
( time timeout 400 $TR --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n8_r02u.json 2> gpurun_out/bench_n8_r02u.err
grep '^{' gpurun_out/bench_n8_r02u.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d.get('slab',{})
print('N=8 value %.4e e2e %.4e ms %.4f'%(d['value'], d['e2e']['value'], d['ms_per_step']))
print('slab', {k:s.get(k) for k in ('ms_per_step','value','speedup_vs_single_gpu','per_rank_kernels_us_max','barriers_and_gaps_us_rank0','error')})
p=s.get('parity') or {}; print('parity', {k:p.get(k) for k in ('q','psi','E','Z','ok')}); print('full', s.get('full_size_check'))"
tail -2 gpurun_out/bench_n8_r02u.err
