#!/bin/bash
# focused re-run: the tests that failed / are new, the coherent-bias measurement, mem_bench
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_runner_gpu.py tests/test_abi.py tests/test_parity_gpu.py tests/test_cfg1_gop_gpu.py -m gpu -q -s > gpurun_out/r2_gputests_b.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2_gputests_b.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2_gputests_b.log | head -20
grep -E "^lanes=|^IP32 free|^device:|seed 1 offset" gpurun_out/r2_gputests_b.log
timeout 300 python tools/acc_bias_coherent.py > gpurun_out/r2_acc_bias_coherent.log 2>&1; cat gpurun_out/r2_acc_bias_coherent.log
timeout 300 python tools/mem_bench.py > gpurun_out/r2_mem1.log 2>&1; tail -20 gpurun_out/r2_mem1.log
