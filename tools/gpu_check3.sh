# conv_hs: kernel tests, micro-bench (one gpurun call)
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "conv_hs" 2>&1 | tail -5
timeout 300 python tools/conv_bench.py hs 2>&1 | tail -18 > gpurun_out/convbench_hs4.log; cat gpurun_out/convbench_hs4.log
for d in 188 191; do echo -n "dbg=$d "; LSSVC_HS_DBG=$d CONV_BENCH_ONLY="3x3 64->64 @1/2" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1; done
LSSVC_HS_DBG=64 CONV_BENCH_ONLY="3x3 64->64 @1/2" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -8
