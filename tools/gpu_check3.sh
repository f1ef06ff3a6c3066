# conv_hs: kernel tests, micro-bench vs conv_h2 (one gpurun call)
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "conv_hs" 2>&1 | tail -25
timeout 300 python tools/conv_bench.py hs h2t 2>&1 | tail -18 > gpurun_out/convbench_hs1.log; cat gpurun_out/convbench_hs1.log
