#!/bin/bash
# developer script (run under gpurun --gpus 2): GPU tests incl. the 2-GPU ones, bench at N=1 and N=2
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,uuid --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi_$TAG.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> gpurun_out/smi_$TAG.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
tail -c 1500 gpurun_out/bench_n1_$TAG.json; tail -5 gpurun_out/bench_n1_$TAG.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err
grep '^{' gpurun_out/bench_n2_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=2 value %.3e e2e %.3e'%(d['value'], d['e2e']['value'])); print(json.dumps(d.get('slab'), indent=1)[:3000])"
tail -8 gpurun_out/bench_n2_$TAG.err
