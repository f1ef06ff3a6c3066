#!/bin/bash
# whole GPU test suite (log in gpurun_out/), then the lanes sweep
timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gputests.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2_gputests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2_gputests.log | head -20
bash tools/gpu_r2_lanes.sh
