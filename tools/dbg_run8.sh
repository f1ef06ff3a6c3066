echo "== compensation off"; LSSVC_ACC_COMP=0 timeout 600 python tools/check_convs.py --size 384 512 --bias --thr 1 > gpurun_out/bias_net_off.log 2>&1; tail -2 gpurun_out/bias_net_off.log
echo "== compensation on"; timeout 600 python tools/check_convs.py --size 384 512 --bias --thr 1 > gpurun_out/bias_net_on.log 2>&1; tail -2 gpurun_out/bias_net_on.log
