# per-pixel-lane gather kernels: kernel tests, A/B against the float4 kernels, 1080p full-size tests
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "not conv_" 2>&1 | tail -3
echo "== new"; timeout 300 python tools/mem_bench.py > gpurun_out/mem4.log 2>&1; cat gpurun_out/mem4.log
echo "== legacy"; LSSVC_GATHER_LEGACY=1 timeout 300 python tools/mem_bench.py 2>&1 | grep -i "warp\|resize"
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s --timeout 600 -x 2>&1 | tail -40
