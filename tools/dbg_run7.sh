CONV_BENCH_BIAS=1 timeout 600 python tools/conv_bench.py hs 2>&1 | grep "\[hs\]" > gpurun_out/bias2.log; cat gpurun_out/bias2.log
timeout 800 python tools/fullsize_parity.py > gpurun_out/fullsize2.log 2>&1; cat gpurun_out/fullsize2.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "conv_hs or conv_pw or ffn or gdn or deconv" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 600 2>&1 | tail -12
