for shape in "study 1x1 128->512" "study 1x1 512->128"; do
export CONV_BENCH_ONLY="$shape"
for d in 0 64 129 130 131 132 136 144 160; do echo "dbg=$d"; LSSVC_HS_DBG=$d timeout 120 python tools/conv_bench.py hs 2>&1 | tail -8 | grep -v "^shape"; done
done > gpurun_out/dbg_hs10.log 2>&1
cat gpurun_out/dbg_hs10.log
