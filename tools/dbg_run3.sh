export CONV_BENCH_ONLY="1x1 64->64 @1"
for d in 0 1 2 4 8 9 16 48 63; do echo "dbg=$d"; LSSVC_H2_DBG=$d timeout 120 python tools/conv_bench.py h2 2>&1 | tail -1; done > gpurun_out/dbg5.log 2>&1
cat gpurun_out/dbg5.log
