#!/bin/bash
# developer script (1 GPU): full GPU test suite, bench, ncu launch list + full capture, large-size parity
TAG=${1:-r02i}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.log
( timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
( timeout 600 python bench.py ) > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err
python - <<PY
import json
for f in ("gpurun_out/bench_n1_$TAG.json", "gpurun_out/bench_default_$TAG.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.4e ms %.4f graph %.4f e2e %.4e step_frac %.3f" % (d["value"], d["ms_per_step"], d["ms_per_step_graph_replay"], d["e2e"]["value"], d["roofline"]["step"]["frac"]),
              {k: (v["us"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "c4", d.get("config4_single_gpu", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 12 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k[1-4]_' -s 24 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/ncu_f_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log
( time timeout 900 python scripts/parity_large.py single 16384 8192 10 ) > gpurun_out/parity_config4_single_$TAG.log 2>&1
tail -5 gpurun_out/parity_config4_single_$TAG.log
( time timeout 1500 python scripts/parity_large.py single 4096 4096 1000 ) > gpurun_out/parity_4096_1000_$TAG.log 2>&1
tail -5 gpurun_out/parity_4096_1000_$TAG.log
