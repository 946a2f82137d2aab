#!/bin/bash
# developer script (run under gpurun): GPU tests, bench, ncu launch list + full capture of each hot kernel
TAG=${1:-r01x}
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 12 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k[1-4]_' -s 24 -c 5 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log; cat gpurun_out/bench_$TAG.json
