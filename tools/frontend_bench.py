"""Front end (lssvc_b200/frontend.py) at 1080p: CUDA-event time of YUV 4:2:0 -> padded RGB, EL -> BL bicubic resize and PSNR,
their algorithmic bytes against the measured copy bandwidth, and the reference's host path (oracle restatement) beside it."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import frontend as fe
from oracle import frontend_oracle as fo

dev = torch.device("cuda:0")
H, W = 1080, 1920
peak = 6544.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rng = np.random.default_rng(0)
frames = [(torch.from_numpy(rng.integers(0, 256, (H, W), dtype=np.uint8)).to(dev), torch.from_numpy(rng.integers(0, 256, (2, H // 2, W // 2), dtype=np.uint8)).to(dev)) for _ in range(4)]
front = fe.FrontEnd(H, W, 2, dev)
Hp, Wp = front.el_size
Hb, Wb = front.bl_size


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


xs = [front.rgb_from_yuv420(*f) for f in frames]
bl = front.base_layer(xs[0])
rows = [
    ("yuv420_to_rgb + pad", timed(lambda i: front.rgb_from_yuv420(*frames[i % 4])), H * W * 1.5 + 3 * Hp * Wp * 4),
    ("imresize cubic x1/2 + clamp (2 passes)", timed(lambda i: front.base_layer(xs[i % 4])), (3 * Hp * Wp + 2 * 3 * Hb * Wp + 3 * Hb * Wb) * 4),
    ("psnr (EL)", timed(lambda i: fe.psnr(xs[i % 4], xs[(i + 1) % 4])), 2 * 3 * Hp * Wp * 4),
]
print(f"{'front end @1080p':42s} {'ms':>8s} {'MB (alg.)':>10s} {'GB/s':>8s}  of {peak:.1f}")
for name, ms, b in rows:
    print(f"{name:42s} {ms:8.4f} {b / 1e6:10.1f} {b / ms / 1e6:8.1f} {100 * b / ms / 1e6 / peak:6.1f}%")
y8, uv8 = frames[0][0].cpu().numpy(), frames[0][1].cpu().numpy()
torch.set_num_threads(os.cpu_count() or 8)
t0 = time.time()
y, uv = fo.read_yuv420_frame(y8.tobytes() + uv8.tobytes(), H, W)
ref_el = fo.pad_el(fo.ycbcr420_to_rgb(y, uv), front.padding["P_HR"])
t1 = time.time()
ref_bl = fo.base_layer(ref_el, front.bl_size)
t2 = time.time()
fo.psnr(ref_el, ref_el.roll(1, 3))
t3 = time.time()
print(f"reference host path (oracle, {torch.get_num_threads()} threads): ycbcr420_to_rgb + pad {1e3 * (t1 - t0):.1f} ms, imresize {1e3 * (t2 - t1):.1f} ms, "
      f"PSNR {1e3 * (t3 - t2):.1f} ms")
