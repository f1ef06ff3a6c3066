# conv_hs: kernel tests, micro-bench (one gpurun call)
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "conv_hs" 2>&1 | tail -5
timeout 300 python tools/conv_bench.py hs 2>&1 | tail -18 > gpurun_out/convbench_hs3.log; cat gpurun_out/convbench_hs3.log
LSSVC_HS_DBG=64 CONV_BENCH_ONLY="3x3 64->64 @1/2" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -8
LSSVC_HS_DBG=64 CONV_BENCH_ONLY="3x3 48->48 @1" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -8
