N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "torchrun rc=$?"; tail -15 gpurun_out/bench_n$N.err | cut -c1-300; head -c 600 gpurun_out/bench_n$N.json; nvidia-smi --query-gpu=index,memory.used --format=csv | head -10; free -g | head -2
