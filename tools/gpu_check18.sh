timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "conv_pw" 2>&1 | tail -2
echo "== double staging"; python tools/pw_bench.py 2>&1 | tail -5
echo "== single staging"; LSSVC_PW_SINGLE_STAGE=1 python tools/pw_bench.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 800 2>&1 | tail -2
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; python -c "import json; d=json.load(open('gpurun_out/bench14.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'])"
LSSVC_PW_SINGLE_STAGE=1 timeout 600 python bench.py --steps 12 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single staging:', d['value'], d['ms_per_step'])"
