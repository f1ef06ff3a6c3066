"""Accumulator-truncation bias of conv_hs for SIGN-COHERENT sums (every product positive) as a function of the number of
accumulation steps T, next to zero-mean data: the data behind ops.acc_comp(steps, coherent=True).
usage (GPU box): python tools/acc_bias_coherent.py > gpurun_out/acc_bias_coherent.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LSSVC_ACC_COMP"] = "0"          # raw bias
import torch
import torch.nn.functional as F

from lssvc_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
print("k  cin  T   signed error (x 2^-24 of mean|y|): zero-mean | positive x, positive w | x^2 (GDN pool), positive w")
for k, cin in ((1, 64), (1, 128), (1, 192), (3, 16), (3, 32), (3, 64), (3, 128), (7, 32)):
    T = k * k * cin // 16
    row = []
    for kind in ("zero", "pos", "square"):
        x = torch.randn(1, cin, 48, 64, generator=g)
        w = torch.randn(64, cin, k, k, generator=g) / (k * k * cin) ** 0.5
        if kind == "pos":
            x, w = x.abs() + 0.5, w.abs()
        if kind == "square":
            x, w = x * x, w.abs()
        pc = ops.PackedConv(w, torch.zeros(64), pad=k // 2, device=dev)
        out = ops.View.alloc(48, 64, 64, dev, zero=True)
        ops.conv(pc, [ops.View.from_nchw(x.to(dev))], out, engine="hs")
        got = out.to_nchw().cpu().double()
        ref = F.conv2d(x.double(), w.double(), None, padding=k // 2)
        row.append(((got - ref) * torch.sign(ref)).mean().item() / ref.abs().mean().item() / 2 ** -24)
    print(f"{k}  {cin:3d} {T:3d}   {row[0]:+8.2f} ({row[0] / T:+.3f} T) | {row[1]:+8.2f} ({row[1] / T:+.3f} T) | {row[2]:+8.2f} ({row[2] / T:+.3f} T)")
