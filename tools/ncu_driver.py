"""One short program that launches every non-conv_hs kernel family of the P-frame at its 1080p size a few times, for
`ncu --set full` (profiles/r2_*_ncu.txt): conv_ffn, conv_pw (plain and with the depthwise front end), dwconv3x3, the gather
/ resample / element-wise kernels and the entropy kernels of tools/mem_bench.py (incl. the group-planar OffsetDiversity).
usage (GPU box): python tools/ncu_driver.py && ncu --set full ... python tools/ncu_driver.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mem_bench  # noqa: E402
from lssvc_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.check(_lib.load().lssvc_device_check(0), "device_check")
H, W = 1152, 1920
g = torch.Generator().manual_seed(0)
# ConvFFN 64 -> 256 -> 64 at full resolution
C, hidden = 64, 256
pf = ops.PackedFfn(torch.randn(hidden, C, 1, 1, generator=g) / math.sqrt(C), torch.randn(hidden, generator=g),
                   torch.randn(C, hidden, 1, 1, generator=g) / math.sqrt(hidden), torch.randn(C, generator=g), dev)
xs = [ops.View(torch.randn(H * W * C, device=dev), H, W, C, C) for _ in range(3)]
out = ops.View.alloc(H, W, C, dev)
for i in range(2):
    ops.ffn(pf, xs[i % 3], out)
# 1x1 64 -> 64 with and without the depthwise 3x3 front end
for dw in (True, False):
    pp = ops.PackedPw(torch.randn(64, 64, 1, 1, generator=g) / 8, torch.randn(64, generator=g), dev,
                      dw_w=torch.randn(64, 1, 3, 3, generator=g) / 3 if dw else None, dw_b=torch.randn(64, generator=g) if dw else None)
    res = ops.View(torch.randn(H * W * 64, device=dev), H, W, 64, 64)
    for i in range(2):
        ops.pw(pp, xs[i % 3], out, act=0.1, res1=res)
# the 1x1 128 -> 512 / 512 -> 128 pair of the C = 128 ConvFFN at 1/4 resolution (general conv kernel)
h4, w4 = H // 4, W // 4
pc1 = ops.PackedConv(torch.randn(512, 128, 1, 1, generator=g) / 11, torch.randn(512, generator=g), pad=0, device=dev)
pc2 = ops.PackedConv(torch.randn(128, 512, 1, 1, generator=g) / 22, torch.randn(128, generator=g), pad=0, device=dev)
x4 = ops.View(torch.randn(h4 * w4 * 128, device=dev), h4, w4, 128, 128)
mid, o4 = ops.View.alloc(h4, w4, 512, dev), ops.View.alloc(h4, w4, 128, dev)
for i in range(2):
    ops.conv(pc1, x4, mid, act=0.1)
    ops.conv(pc2, mid, o4, act=0.1, res1=x4)
# the headline layer: 3x3 64 -> 64 at half resolution
h2, w2 = H // 2, W // 2
pc3 = ops.PackedConv(torch.randn(64, 64, 3, 3, generator=g) / 24, torch.zeros(64), device=dev)
x2 = [ops.View(torch.randn(h2 * w2 * 64, device=dev), h2, w2, 64, 64) for _ in range(2)]
o2 = ops.View.alloc(h2, w2, 64, dev)
for i in range(2):
    ops.conv(pc3, x2[i], o2, act=0.01)
torch.cuda.synchronize()
mem_bench.WARM, mem_bench.REPS = 1, 1
rows = mem_bench.run(torch, dev)
for r in rows:
    print(f"{r['kernel']:44s} {r['ms']:8.4f} ms {r['gbs']:8.1f} GB/s")
torch.cuda.synchronize()
print("ncu_driver done")
