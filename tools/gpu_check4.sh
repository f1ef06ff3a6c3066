for v in "LSSVC_HS_MT=1" "LSSVC_HS_MT=2" "LSSVC_HS_MT=1 LSSVC_H2_NOTMA=1" "LSSVC_HS_MT=2 LSSVC_H2_NOTMA=1"; do
  echo "== $v"; env $v timeout 300 python tools/conv_bench.py hs 2>&1 | tail -15
done > gpurun_out/convbench_hs2.log 2>&1
cat gpurun_out/convbench_hs2.log
