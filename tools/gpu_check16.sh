timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_ratio_gpu.py -m gpu -q --timeout 800 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; tail -1 gpurun_out/bench12.err; cut -c1-200 gpurun_out/bench12.json; python -c "import json; d=json.load(open('gpurun_out/bench12.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
