"""The kernels added late in round 2, at their 1080p P-frame sizes, for `ncu --set full` (profiles/r2_head_entropy_ncu.txt):
conv_head (3x3 64->2, 3x3 48->3, 7x7 16->2 at full resolution) and conv_hs with the entropy epilogues (BitEstimator on the 3x3
stride-2 hyper-encoder tail, Laplace on a 2 x 64-channel parameter conv, Laplace over two channel tiles (2 x 96), four-part on
the 1x1 1024 -> 256 ConvFFN tail with LeakyReLU + residual).
usage (GPU box): python tools/ncu_driver2.py && ncu --set full -k regex:'conv_head|conv_hs' ... python tools/ncu_driver2.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import _lib, entropy, ops  # noqa: E402

dev = torch.device("cuda:0")
_lib.check(_lib.load().lssvc_device_check(0), "device_check")
g = torch.Generator().manual_seed(0)
H, W = 1152, 1920
for cin, cout, k in ((64, 2, 3), (48, 3, 3), (16, 2, 7)):
    pc = ops.PackedConv(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k), torch.randn(cout, generator=g), pad=k // 2, device=dev)
    x = ops.View(torch.randn(H * W * cin, device=dev), H, W, cin, cin)
    out = ops.View.alloc(H, W, 8, dev, zero=True).slice(0, cout)
    for _ in range(2):
        ops.conv(pc, x, out)
thr = entropy.video_scale_thresholds().to(dev)
bits = torch.zeros(1, dtype=torch.float64, device=dev)
h, w = H // 16, W // 16
# BitEstimator epilogue: hyper-encoder tail 3x3 stride 2, 64 -> 64 (z at 1/64 resolution)
coef = torch.cat([torch.rand(4, 64, generator=g) + 0.5, torch.randn(4, 64, generator=g) * 0.5, torch.tanh(torch.randn(3, 64, generator=g))], 0).t().contiguous().to(dev)
pc = ops.PackedConv(torch.randn(64, 64, 3, 3, generator=g) / 24, torch.randn(64, generator=g), stride=2, pad=1, device=dev)
x = ops.View(torch.randn((h // 2) * (w // 2) * 64, device=dev) * 3, h // 2, w // 2, 64, 64)
for _ in range(2):
    ops.conv(pc, x, ops.View.alloc(h // 4, w // 4, 64, dev), entropy={"mode": "bitparm", "coef": coef, "bits": bits})
# Laplace epilogues: 3x3 128 -> 2 x 64 (EL mv_y, one tile) and 1x1 320 -> 2 x 96 at the BL's 1/16 (two tiles of 96, interleaved)
for cin, C, k, hh, ww in ((128, 64, 3, h, w), (320, 96, 1, h // 2, w // 2)):
    wt = torch.randn(2 * C, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    pc = ops.PackedConv(wt, torch.randn(2 * C, generator=g), pad=k // 2, device=dev, pair_tile=ops.laplace_pair_tile(2 * C))
    x = ops.View(torch.randn(hh * ww * cin, device=dev), hh, ww, cin, cin)
    y = ops.View(torch.randn(hh * ww * C, device=dev) * 4, hh, ww, C, C)
    for _ in range(2):
        ops.conv(pc, x, ops.View.alloc(hh, ww, 2 * C, dev), entropy={"mode": "laplace", "y": y, "y_hat": ops.View.alloc(hh, ww, C, dev), "bits": bits,
                                                                   "sym": torch.zeros(C * hh * ww, dtype=torch.int32, device=dev),
                                                                   "index": torch.zeros(C * hh * ww, dtype=torch.int32, device=dev), "thresholds": thr})
# four-part epilogue: ConvFFN tail 1x1 1024 -> 256 + LeakyReLU + residual, step 1 of 4
C = 128
pc = ops.PackedConv(torch.randn(2 * C, 1024, 1, 1, generator=g) / 32, torch.randn(2 * C, generator=g), pad=0, device=dev, pair_tile=ops.laplace_pair_tile(2 * C))
x = ops.View(torch.randn(h * w * 1024, device=dev), h, w, 1024, 1024)
r = ops.View(torch.randn(h * w * 2 * C, device=dev), h, w, 2 * C, 2 * C)
y = ops.View(torch.randn(h * w * C, device=dev) * 4, h, w, C, C)
y_hat = ops.View.alloc(h, w, C, dev, zero=True)
for _ in range(2):
    ops.conv(pc, x, ops.View.alloc(h, w, 2 * C, dev), act=0.1, res1=r, entropy={"mode": "fourpart", "step": 1, "y": y, "y_hat": y_hat, "bits": bits, "thresholds": thr})
torch.cuda.synchronize()
print("ncu_driver2 done", bits.item())
