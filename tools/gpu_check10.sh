timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "not conv_" 2>&1 | tail -4
timeout 300 python tools/mem_bench.py > gpurun_out/mem2.log 2>&1; cat gpurun_out/mem2.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
