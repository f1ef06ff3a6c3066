for shape in "study 1x1 128->512" "study 1x1 512->128" "1x1 64->256 @1" "1x1 256->64 @1" "3x3 128->256ps" "3x3 192->192"; do
export CONV_BENCH_ONLY="$shape"
for nt in 128 64 32; do echo "ntile=$nt"; LSSVC_HS_NTILE=$nt timeout 120 python tools/conv_bench.py hs 2>&1 | tail -1; done
done > gpurun_out/dbg_hs11.log 2>&1
cat gpurun_out/dbg_hs11.log
