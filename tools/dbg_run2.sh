export CONV_BENCH_ONLY="3x3 64->64 @1/2"
timeout 120 python tools/conv_bench.py h2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_h2 -s 5 -c 1 -o gpurun_out/prof_h2_r1e python tools/conv_bench.py h2 > gpurun_out/ncu_h2.log 2>&1; tail -2 gpurun_out/ncu_h2.log
