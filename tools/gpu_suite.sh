#!/bin/bash
# One parameterised entry point for the GPU box:  tools/gpurun_retry.sh <timeout> 'bash tools/gpu_suite.sh <what> ...'
#   tests      whole `-m gpu` suite                     -> gpurun_out/<tag>_gputests.log
#   bench      bench.py (default arguments)             -> gpurun_out/<tag>_bench.json
#   lanes      bench.py --lanes 1 / 2 / 3               -> gpurun_out/<tag>_bench_l<k>.json
#   streams    tools/stream_bench.py, both stream paths -> gpurun_out/<tag>_stream_*.log
#   mem        tools/mem_bench.py                       -> gpurun_out/<tag>_mem.log
#   layers     tools/layer_times.py (per-layer table)   -> gpurun_out/<tag>_layers.log
#   launches   ncu launch list of bench.py              -> gpurun_out/<tag>_launches.csv
#   ncu        ncu --set full of tools/ncu_driver.py    -> gpurun_out/<tag>_kernels_raw.csv (raw metrics page)
#   ffn        tools/ffn_bench.py + pw_bench.py         -> gpurun_out/<tag>_ffn.log
#   ncu2       ncu --set full of tools/ncu_driver2.py   -> gpurun_out/<tag>_head_entropy.ncu-rep (conv_head, entropy epilogues)
# TAG=<tag> (default r2) names the outputs.
TAG=${TAG:-r2}
for what in "$@"; do
  case $what in
    tests) timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"
           grep -E "passed|failed" gpurun_out/${TAG}_gputests.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_gputests.log | head -20;;
    bench) timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/${TAG}_bench.json;;
    lanes) for k in 1 2 3; do timeout 600 python bench.py --lanes $k --steps 24 --no-cpu-baseline > gpurun_out/${TAG}_bench_l$k.json 2> gpurun_out/${TAG}_bench_l$k.err
             python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_l$k.json'));print($k,d['value'],d['e2e']['value'],d['hbm_gb'])"; done;;
    streams) timeout 300 python tools/stream_bench.py --p-frames 5 > gpurun_out/${TAG}_stream_genuine.log 2>&1; cat gpurun_out/${TAG}_stream_genuine.log
             timeout 300 python tools/stream_bench.py --p-frames 5 --single-pass > gpurun_out/${TAG}_stream_single.log 2>&1; cat gpurun_out/${TAG}_stream_single.log;;
    ffn) (timeout 300 python tools/ffn_bench.py; timeout 300 python tools/pw_bench.py) > gpurun_out/${TAG}_ffn.log 2>&1; cat gpurun_out/${TAG}_ffn.log;;
    mem) timeout 300 python tools/mem_bench.py > gpurun_out/${TAG}_mem.log 2>&1; cat gpurun_out/${TAG}_mem.log;;
    layers) timeout 300 python tools/layer_times.py > gpurun_out/${TAG}_layers.log 2>&1; head -70 gpurun_out/${TAG}_layers.log;;
    launches) python bench.py --steps 6 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
              ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
                  python bench.py --steps 6 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?";;
    ncu) # ~60 matched launches at ~5 s each under --set full; raw metrics as CSV (a .ncu-rep of this run exceeds the 64 MiB gpurun_out limit)
         python tools/ncu_driver.py > gpurun_out/${TAG}_ncu_plain.log 2>&1 &&
         ncu --set full --clock-control none -k regex:'conv_ffn|conv_pw|conv_hs|dwconv|od_|offset_div|flow_warp|laplace|four_part|bilinear|nhwc|pool2|softmax2|lrelu_copy' \
             -c 80 --csv --page raw --log-file gpurun_out/${TAG}_kernels_raw.csv python tools/ncu_driver.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
         ls -la gpurun_out/${TAG}_kernels_raw.csv;;
    ncu2) # conv_head + the entropy epilogues of conv_hs (tools/ncu_driver2.py); the .ncu-rep is small enough to travel
         python tools/ncu_driver2.py > gpurun_out/${TAG}_ncu2_plain.log 2>&1 &&
         ncu --set full --clock-control none --import-source on -k regex:'conv_head|conv_hs' -c 16 -f -o gpurun_out/${TAG}_head_entropy \
             python tools/ncu_driver2.py > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -2 gpurun_out/${TAG}_ncu2.log; ls -la gpurun_out/${TAG}_head_entropy.ncu-rep;;
    *) echo "unknown: $what";;
  esac
done
