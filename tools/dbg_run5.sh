# conv_hs operand-fetch study: MMA-only timings with the A tap windows / 8-row group stride forced onto 1024-byte swizzle atoms
export CONV_BENCH_ONLY="3x3 64->64 @1/2"
for d in 188 8380 16572 24764 32956 57532 189 24765 190 24766 128 32896; do echo "dbg=$d"; LSSVC_HS_DBG=$d timeout 120 python tools/conv_bench.py hs 2>&1 | tail -1; done > gpurun_out/dbg_hs9.log 2>&1
cat gpurun_out/dbg_hs9.log
