python tools/pw_bench.py > gpurun_out/pw1.log 2>&1; cat gpurun_out/pw1.log
PW_BENCH_ONLY=64x64x1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pw -s 4 -c 1 -f -o gpurun_out/prof_pw_r1 python tools/pw_bench.py > gpurun_out/ncu_pw.log 2>&1; tail -2 gpurun_out/ncu_pw.log
