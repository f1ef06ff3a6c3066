timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x -k "h2" 2>&1 | tail -2
export CONV_BENCH_ONLY="3x3 64->64 @1/2"
for d in 0 1 8 9 63; do echo "dbg=$d"; LSSVC_H2_DBG=$d timeout 120 python tools/conv_bench.py h2 2>&1 | tail -1; done > gpurun_out/dbg4.log 2>&1
cat gpurun_out/dbg4.log
unset CONV_BENCH_ONLY
timeout 300 python tools/conv_bench.py h2 2>&1 | tail -30 > gpurun_out/convbench5.log; cat gpurun_out/convbench5.log
