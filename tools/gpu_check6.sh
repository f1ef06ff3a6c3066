# full check: kernel tests, parity, layer times, bench
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -s 2>&1 | grep -E "frame \(|teacher|passed|failed|Error|assert" | cut -c1-200 | tail -30
timeout 600 python tools/layer_times.py > gpurun_out/layers4.log 2>&1; head -22 gpurun_out/layers4.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -3 gpurun_out/bench5.err; cat gpurun_out/bench5.json
