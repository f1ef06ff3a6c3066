python tools/mem_bench.py > gpurun_out/plain_mem.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"flow_warp4|bilinear_resize4|dwconv3x3_x2|offset_diversity" -s 12 -c 4 -f -o gpurun_out/prof_mem_r1 python tools/mem_bench.py > gpurun_out/ncu_mem.log 2>&1
tail -3 gpurun_out/ncu_mem.log
