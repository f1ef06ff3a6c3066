timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "conv_hs or gdn or deconv" 2>&1 | tail -4
timeout 300 python tools/conv_bench.py hs 2>&1 | tail -16
for ONLY in "3x3 64->64 @1/2" "3x3 48->48 @1"; do
for x in r rs; do
  echo -n "extras=$x  "; CONV_BENCH_EXTRAS=$x CONV_BENCH_ONLY="$ONLY" timeout 100 python tools/conv_bench.py hs 2>&1 | tail -1
done
done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
