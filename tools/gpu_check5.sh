ONLY="${ONLY:-3x3 64->64 @1/2}"
echo "== plain"; CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1
for d in 64 127; do
  echo "== dbg=$d"; LSSVC_HS_DBG=$d CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -8
done
