# round deliverables in one call: smoke, full GPU suite, default bench, reference arm, ncu launch list
set -x
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -2 gpurun_out/bench7.err; cat gpurun_out/bench7.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err; cat gpurun_out/bench7_ref.json
timeout 300 python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_frame.py --size 1080p --p-frames 2 > gpurun_out/ncu_frame.log 2>&1
tail -3 gpurun_out/ncu_frame.log
