#!/bin/bash
# runner test + lanes sweep of bench.py (gpurun -- 'bash tools/gpu_r2_lanes.sh')
timeout 600 python -m pytest tests/test_runner_gpu.py -m gpu -x -q -s > gpurun_out/r2_runner.log 2>&1; tail -5 gpurun_out/r2_runner.log
for cfg in "1 --no-graphs" "1" "2" "3"; do
  set -- $cfg
  timeout 600 python bench.py --lanes $1 $2 --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_l$1$2.json 2> gpurun_out/r2_bench_l$1$2.err
  echo "lanes $cfg rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_l$1$2.json").read())
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "hbm_gb", "clocks")}, d["e2e"]["value"], d["roofline"]["in_frame"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/r2_bench_l$1$2.err").read()[-1500:])
PY
done
