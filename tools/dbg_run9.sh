timeout 800 python tools/fullsize_parity.py > gpurun_out/fullsize3.log 2>&1; grep "h2\|x_hat" gpurun_out/fullsize3.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s --timeout 800 2>&1 | grep -v "^  " | tail -30
timeout 600 python tools/check_convs.py --size 384 512 --bias --thr 1 > gpurun_out/bias_net_on2.log 2>&1; tail -2 gpurun_out/bias_net_on2.log
