timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x -k "h2 or gdn" 2>&1 | tail -5
timeout 300 python tools/conv_bench.py h2 2>&1 | tail -30 > gpurun_out/convbench7.log; cat gpurun_out/convbench7.log
