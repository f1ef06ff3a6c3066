#!/bin/bash
# whole GPU suite, stream timing (both paths), mem_bench, bench lanes=1
timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gputests_c.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2_gputests_c.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2_gputests_c.log | head -20
grep -E "^IP32 free|symbols:|^4K|^1080p|single pass" gpurun_out/r2_gputests_c.log | head -40
timeout 300 python tools/stream_bench.py --p-frames 5 > gpurun_out/r2_stream_genuine.log 2>&1; cat gpurun_out/r2_stream_genuine.log
timeout 300 python tools/stream_bench.py --p-frames 5 --single-pass > gpurun_out/r2_stream_single.log 2>&1; cat gpurun_out/r2_stream_single.log
timeout 300 python tools/mem_bench.py > gpurun_out/r2_mem2.log 2>&1; grep -E "offset|nchw" gpurun_out/r2_mem2.log
timeout 600 python bench.py --lanes 1 --no-graphs --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c.json").read())
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["in_frame"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/r2_bench_c.err").read()[-1500:])
PY
