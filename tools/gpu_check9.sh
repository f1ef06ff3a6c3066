timeout 600 python tools/layer_times.py --top 70 > gpurun_out/layers8.log 2>&1; head -40 gpurun_out/layers8.log
