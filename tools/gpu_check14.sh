timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "not (conv_tc or conv_simt or conv_h2)" 2>&1 | tail -3
timeout 300 python tools/mem_bench.py > gpurun_out/mem5.log 2>&1; head -6 gpurun_out/mem5.log
CONV_BENCH_ONLY="3x3 64->64 @1/2" timeout 200 python tools/conv_bench.py hs 2>&1 | tail -1
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s --timeout 800 > gpurun_out/parity16.log 2>&1; tail -3 gpurun_out/parity16.log; grep -A12 "^1080p\|\.1080p" gpurun_out/parity16.log | head -60
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -2 gpurun_out/bench10.err; cut -c1-400 gpurun_out/bench10.json
timeout 600 python tools/layer_times.py --top 60 > gpurun_out/layers13.log 2>&1; head -20 gpurun_out/layers13.log
