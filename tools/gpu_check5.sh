# 188 = MMAs only; +4096: A_lo MMAs accumulate into the other accumulator set (independent of the A_hi chain)
for ONLY in "3x3 64->64 @1/2" "study 3x3 64->32" "study 3x3 64->16"; do
for d in 188 4284; do
  echo -n "dbg=$d  "; LSSVC_HS_DBG=$d CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1
done
done
