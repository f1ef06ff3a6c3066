set -x
export CONV_BENCH_ONLY="3x3 64->64 @1/2"
python tools/conv_bench.py hs > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_hs -s 6 -c 2 -f -o gpurun_out/prof_hs_r1g python tools/conv_bench.py hs > gpurun_out/ncu_hs.log 2>&1
tail -2 gpurun_out/plain.log; tail -2 gpurun_out/ncu_hs.log
unset CONV_BENCH_ONLY
timeout 300 python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/ncu_frame.log 2>&1
tail -3 gpurun_out/ncu_frame.log
