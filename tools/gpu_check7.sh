timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x -k "graph" -s 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -3 gpurun_out/bench6.err; python -c "
import json;d=json.load(open('gpurun_out/bench6.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','conv_tflops','clocks')}, d['roofline']['achieved'])"
