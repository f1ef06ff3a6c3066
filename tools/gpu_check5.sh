for ONLY in "3x3 64->64 @1/2" "3x3 48->48 @1"; do
for x in "" l r rs ro rso; do
  echo -n "extras=$x  "; CONV_BENCH_EXTRAS=$x CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1
done
done
