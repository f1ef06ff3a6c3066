# multi-GPU bench (2 ranks) + 4K smoke of one I + 2 P frames
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 600 python tools/profile_frame.py --size 4k --p-frames 2 2>&1 | tail -4
nvidia-smi --query-gpu=memory.used --format=csv | tail -2
