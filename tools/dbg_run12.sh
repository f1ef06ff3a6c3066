LSSVC_FFN_DBG=1 timeout 200 python tools/ffn_bench.py > gpurun_out/ffn_dbg1.log 2>&1; grep -A7 "conv_ffn prof" gpurun_out/ffn_dbg1.log | head -40
timeout 200 python tools/ffn_bench.py 2>&1 | tail -4
