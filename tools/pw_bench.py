"""Resident-weight 1x1 conv with the fused depthwise 3x3 front end (csrc/conv_pw.cu) at the sizes of the 1080p P-frame:
CUDA-event time, GB/s on the in + out minimum.  usage: python tools/pw_bench.py"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import _lib, ops

dev = torch.device("cuda:0")
_lib.check(_lib.load().lssvc_device_check(0), "device_check")
only = os.environ.get("PW_BENCH_ONLY")
for cin, cout, H, W, dw in ((64, 64, 1152, 1920, True), (64, 48, 1152, 1920, True), (48, 32, 1152, 1920, True), (32, 64, 576, 960, True), (64, 64, 1152, 1920, False)):
    if only and only != f"{cin}x{cout}x{int(dw)}":
        continue
    g = torch.Generator().manual_seed(0)
    w = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin); b = torch.randn(cout, generator=g)
    dw_w = torch.randn(cin, 1, 3, 3, generator=g) / 3 if dw else None
    dw_b = torch.randn(cin, generator=g) if dw else None
    pp = ops.PackedPw(w, b, dev, dw_w=dw_w, dw_b=dw_b)
    xs = [ops.View(torch.randn(H * W * cin, device=dev), H, W, cin, cin) for _ in range(3)]
    res = ops.View(torch.randn(H * W * cout, device=dev), H, W, cout, cout)
    out = ops.View.alloc(H, W, cout, dev)
    fn = lambda i: ops.pw(pp, xs[i % 3], out, act=0.1, res1=res)
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(9): fn(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 9
    gb = H * W * (cin + 2 * cout) * 4 / 1e9
    print(f"pw {cin}->{cout} dw={int(dw)} {H}x{W} (+res1): {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s of the in + res + out minimum)", flush=True)
