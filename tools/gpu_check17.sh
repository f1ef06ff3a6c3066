timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "conv_pw or ffn or conv_hs" 2>&1 | tail -2
python tools/pw_bench.py 2>&1 | tail -5
python tools/ffn_bench.py 2>&1 | tail -4
CONV_BENCH_ONLY="3x3 64->64 @1/2" timeout 200 python tools/conv_bench.py hs 2>&1 | tail -1
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench13.json 2> gpurun_out/bench13.err; python -c "import json; d=json.load(open('gpurun_out/bench13.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'])"
