# kernel tests, parity tests, conv micro-bench, bench (one gpurun call)
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x 2>&1 | tail -6
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -s 2>&1 | grep -E "frame \(|teacher|passed|failed|Error|assert" | cut -c1-160 | grep -E "h2|passed|failed|Error|assert" 
timeout 300 python tools/conv_bench.py h2 2>&1 | tail -16 > gpurun_out/convbench6.log; cat gpurun_out/convbench6.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; python -c "
import json;d=json.load(open('gpurun_out/bench3.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','conv_tflops','clocks')}, d['roofline']['achieved'])"
