timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "ffn" 2>&1 | tail -3
timeout 600 python tools/layer_times.py --top 4 > gpurun_out/layers11.log 2>&1; head -26 gpurun_out/layers11.log | grep -E "wall|ffn|conv\[|pw"
