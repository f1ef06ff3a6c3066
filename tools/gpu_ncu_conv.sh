# ncu --set full of the conv kernel on one shape (CONV_BENCH_ONLY), after a plain run of the same command
export CONV_BENCH_ONLY="${ONLY:-3x3 64->64 @1/2}"
python tools/conv_bench.py hs > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_hs -s 6 -c 2 -f -o gpurun_out/prof_hs_${TAG:-r1f} python tools/conv_bench.py hs > gpurun_out/ncu_hs.log 2>&1
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_hs.log
