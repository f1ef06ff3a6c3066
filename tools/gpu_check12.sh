# after the hi0 hi1 lo0 lo1 issue order: conv_hs tests, frame parity, bench
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "conv_hs" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -2 gpurun_out/bench9.err; cat gpurun_out/bench9.json
