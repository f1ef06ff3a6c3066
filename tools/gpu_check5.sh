for mt in 1 2; do
for ONLY in "study 3x3 64->16" "study 3x3 64->32" "3x3 64->64 @1/2" "study 3x3 64->96" "study 3x3 64->128"; do
for d in 188 190 189; do
  echo -n "mt=$mt dbg=$d  "; LSSVC_HS_MT=$mt LSSVC_HS_DBG=$d CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1
done
done
done
