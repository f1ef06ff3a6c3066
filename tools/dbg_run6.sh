export CONV_BENCH_BIAS=1
timeout 600 python tools/conv_bench.py hs simt 2>&1 | grep -v "^3x3\|^7x7\|^1x1\|^shape" > gpurun_out/bias1.log; cat gpurun_out/bias1.log
