"""Fused ConvFFN (csrc/conv_ffn.cu) vs the two 1x1 conv launches it replaces, CUDA-event time at 1080p sizes."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import _lib, ops

dev = torch.device("cuda:0")
_lib.check(_lib.load().lssvc_device_check(0), "device_check")
for C, hidden, H, W in ((64, 256, 1152, 1920), (48, 192, 1152, 1920), (32, 128, 1152, 1920), (64, 256, 576, 960)):
    g = torch.Generator().manual_seed(0)
    w1 = torch.randn(hidden, C, 1, 1, generator=g) / math.sqrt(C); b1 = torch.randn(hidden, generator=g)
    w2 = torch.randn(C, hidden, 1, 1, generator=g) / math.sqrt(hidden); b2 = torch.randn(C, generator=g)
    pf = ops.PackedFfn(w1, b1, w2, b2, dev)
    pc1 = ops.PackedConv(w1, b1, pad=0, device=dev); pc2 = ops.PackedConv(w2, b2, pad=0, device=dev)
    xs = [ops.View(torch.randn(H * W * C, device=dev), H, W, C, C) for _ in range(3)]
    out = ops.View.alloc(H, W, C, dev); mid = ops.View.alloc(H, W, hidden, dev)
    def fused(i): ops.ffn(pf, xs[i % 3], out)
    def unfused(i):
        ops.conv(pc1, xs[i % 3], mid, act=0.1)
        ops.conv(pc2, mid, out, act=0.1, res1=xs[i % 3])
    res = {}
    for name, fn in (("fused", fused), ("unfused", unfused)):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(9): fn(i)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 9
    gb = 2.0 * H * W * C * 4 / 1e9
    fl = 4.0 * H * W * C * hidden
    print(f"C={C} hidden={hidden} {H}x{W}: fused {res['fused']:.3f} ms ({gb / res['fused'] * 1e3:.0f} GB/s of the 2*C*4 B/px minimum, "
          f"{fl / res['fused'] / 1e9:.0f} TFLOP/s)   unfused {res['unfused']:.3f} ms", flush=True)
