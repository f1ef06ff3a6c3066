timeout 600 python tools/layer_times.py --top 12 > gpurun_out/layers7.log 2>&1; head -34 gpurun_out/layers7.log
python - <<'PY'
import time, torch, sys, os
sys.path.insert(0, os.getcwd())
import bench
from lssvc_b200 import _lib
dev = torch.device("cuda:0")
frames, shape_hr = bench.make_frames(bench.SIZES["1080p"], 3, seed=0)
coder = bench.Coder(dev, shape_hr)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    coder.step(0, 1, frames[0][0].to(dev), frames[0][1].to(dev))
    torch.cuda.synchronize(); print(f"I-frame {rep}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
PY
