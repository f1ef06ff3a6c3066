# conv_pw kernel test, layer times, parity (one gpurun call)
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "conv_pw" -s 2>&1 | tail -12
timeout 600 python tools/layer_times.py > gpurun_out/layers3.log 2>&1; head -20 gpurun_out/layers3.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
