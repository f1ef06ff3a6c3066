"""Narrow-head convolutions (Cout 2..3): the register-blocked fp32 kernel (csrc/conv_head.cu) against the tensor-core kernel
(conv_hs) on the head shapes of the 1080p P-frame; CUDA-event time per launch, rotating over 3 input buffers.

usage: python tools/head_bench.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import ops  # noqa: E402

SHAPES = [("3x3 64->2 @1", 64, 2, 3, 1152, 1920), ("3x3 48->3 @1", 48, 3, 3, 1152, 1920), ("7x7 16->2 @1", 16, 2, 7, 1152, 1920),
          ("3x3 64->3 @1/2", 64, 3, 3, 576, 960), ("3x3 64->2 @1/2", 64, 2, 3, 576, 960), ("7x7 16->2 @1/2", 16, 2, 7, 576, 960),
          ("7x7 16->2 @1/4", 16, 2, 7, 288, 480), ("7x7 16->2 @1/16", 16, 2, 7, 72, 120)]


def main():
    dev = torch.device("cuda:0")
    for name, cin, cout, k, H, W in SHAPES:
        g = torch.Generator().manual_seed(1)
        w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
        pc = ops.PackedConv(w, torch.randn(cout, generator=g), stride=1, pad=k // 2, device=dev)
        bufs = [ops.View(torch.randn(H * W * cin, device=dev), H, W, cin, cin) for _ in range(3)]
        out = ops.View.alloc(H, W, 8, dev, zero=True).slice(0, cout)
        row = []
        for eng in (None, "hs"):
            for i in range(3):
                ops.conv(pc, bufs[i], out, engine=eng)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 30
            e0.record()
            for i in range(n):
                ops.conv(pc, bufs[i % 3], out, engine=eng)
            e1.record()
            torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) / n)
        fma = H * W * k * k * cin * cout
        print(f"{name:18s} head {row[0]:.3f} ms ({fma / row[0] / 1e9:6.2f} TFMA/s fp32, {H * W * cin * 4 / row[0] / 1e6:6.0f} GB/s in)   conv_hs {row[1]:.3f} ms")


if __name__ == "__main__":
    main()
