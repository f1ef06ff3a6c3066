set -x
timeout 1500 python -m pytest tests/ -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -1 gpurun_out/bench11.err; cut -c1-300 gpurun_out/bench11.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench11_ref.json 2> gpurun_out/bench11_ref.err; tail -1 gpurun_out/bench11_ref.err; cut -c1-400 gpurun_out/bench11_ref.json
# ncu: launch list of two P-frames, then full captures of the headline conv and of the HBM-bound kernels
timeout 300 python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/plain3.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/ncu_frame2.log 2>&1
tail -3 gpurun_out/ncu_frame2.log
export CONV_BENCH_ONLY="3x3 64->64 @1/2"
python tools/conv_bench.py hs > gpurun_out/plain4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_hs -s 6 -c 1 -f -o gpurun_out/prof_hs_r1h python tools/conv_bench.py hs > gpurun_out/ncu_hs2.log 2>&1
unset CONV_BENCH_ONLY
python tools/mem_bench.py > gpurun_out/mem6.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flow_warp_q|bilinear_up2_q|dwconv3x3_x2|offset_diversity|softmax2_blend4|laplace_quant" -c 24 -f -o gpurun_out/prof_mem_r1b python tools/mem_bench.py > gpurun_out/ncu_mem2.log 2>&1
tail -2 gpurun_out/ncu_mem2.log
