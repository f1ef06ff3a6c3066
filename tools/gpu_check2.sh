timeout 600 python tools/layer_times.py --top 8 > gpurun_out/layers10.log 2>&1; head -24 gpurun_out/layers10.log
LSSVC_FUSE_PW=0 timeout 600 python tools/layer_times.py --top 8 2>&1 | head -8
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
