for shape in "study 1x1 128->512" "study 1x1 512->128" "1x1 256->64 @1" "1x1 64->256 @1"; do
export CONV_BENCH_ONLY="$shape"
for cfg in "128 0" "64 1" "64 2" "32 1" "32 2"; do set -- $cfg; echo "ntile=$1 mt=$2"; LSSVC_HS_NTILE=$1 LSSVC_HS_MT=$2 timeout 120 python tools/conv_bench.py hs 2>&1 | tail -1; done
done > gpurun_out/dbg_hs13.log 2>&1
cat gpurun_out/dbg_hs13.log
