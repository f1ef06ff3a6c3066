timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err; python -c "
import json;d=json.load(open('gpurun_out/bench8.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','conv_tflops','clocks')}, d['roofline']['achieved'], d['roofline']['frac'])"
timeout 300 python tools/conv_bench.py hs 2>&1 | tail -16 > gpurun_out/convbench_hs5.log; cat gpurun_out/convbench_hs5.log
timeout 300 python tools/mem_bench.py > gpurun_out/mem3.log 2>&1; cat gpurun_out/mem3.log
