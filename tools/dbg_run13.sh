export CONV_BENCH_ONLY="study 1x1 128->512"
for cfg in "128 2 6" "64 1 2" "64 1 3" "64 1 4" "64 1 6" "64 2 2" "64 2 3"; do set -- $cfg; echo "ntile=$1 mt=$2 halos<=$3"; LSSVC_HS_DBG=64 LSSVC_HS_NTILE=$1 LSSVC_HS_MT=$2 LSSVC_HS_HALOS=$3 timeout 120 python tools/conv_bench.py hs 2>&1 | grep -E "conv_hs prof|mma  |study" | tail -3; done > gpurun_out/dbg_hs12.log 2>&1
cat gpurun_out/dbg_hs12.log
