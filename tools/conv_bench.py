"""GPU micro-benchmark + accuracy check of the convolution engines on the layer shapes that dominate the 1080p P-frame
(SURVEY.md App. B).  For each shape: max error relative to the output scale against an fp64 torch conv (on a reduced
height so the reference stays cheap) and CUDA-event time per launch at the full size, rotating over 3 input buffers.

usage: python tools/conv_bench.py [engine ...]      (default: h2 simt)
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import _lib, ops  # noqa: E402

# name, src channels, cout, k, stride, H, W, pixel shuffle
SHAPES = [
    ("3x3 64->64 @1/2", [64], 64, 3, 1, 576, 960, False),
    ("3x3 64->64 @1", [64], 64, 3, 1, 1152, 1920, False),
    ("3x3 48->48 @1", [48], 48, 3, 1, 1152, 1920, False),
    ("3x3 64+64->64 @1/2", [64, 64], 64, 3, 1, 576, 960, False),
    ("3x3 96->64 @1/4", [96], 64, 3, 1, 288, 480, False),
    ("3x3 128->256ps @1/4", [128], 256, 3, 1, 288, 480, True),
    ("3x3 s2 64->64 @1->1/2", [64], 64, 3, 2, 1152, 1920, False),
    ("7x7 32->64 @1/2", [32], 64, 7, 1, 576, 960, False),
    ("7x7 64->32 @1", [64], 32, 7, 1, 1152, 1920, False),
    ("1x1 64->64 @1", [64], 64, 1, 1, 1152, 1920, False),
    ("1x1 64->256 @1", [64], 256, 1, 1, 1152, 1920, False),
    ("1x1 256->64 @1", [256], 64, 1, 1, 1152, 1920, False),
    ("3x3 192->192 @1/8(BL)", [192], 192, 3, 1, 144, 240, False),
    ("3x3 64->3 @1", [64], 3, 3, 1, 1152, 1920, False),
    ("3x3 128->128 @1/16", [128], 128, 3, 1, 72, 120, False),
]
# extra shapes for kernel studies, selected with CONV_BENCH_ONLY (not part of the default table)
EXTRA = [
    ("study 3x3 64->128 @1/2", [64], 128, 3, 1, 576, 960, False),
    ("study 3x3 64->32 @1/2", [64], 32, 3, 1, 576, 960, False),
    ("study 3x3 64->16 @1/2", [64], 16, 3, 1, 576, 960, False),
    ("study 3x3 64->96 @1/2", [64], 96, 3, 1, 576, 960, False),
    ("study 1x1 128->512 @1/4", [128], 512, 1, 1, 288, 480, False),
    ("study 1x1 512->128 @1/4", [512], 128, 1, 1, 288, 480, False),
]


def run(engine, shape, dev, check_rows=64):
    name, cins, cout, k, stride, H, W, ps = shape
    g = torch.Generator().manual_seed(1)
    cin = sum(cins)
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    if os.environ.get("CONV_BENCH_BIAS"):        # accumulation-bias study: no conv bias, signed error statistics printed
        b = torch.zeros(cout)
    pad = k // 2 if k > 1 else 0
    pc = ops.PackedConv(w, b, stride=stride, pad=pad, src_channels=[(c, c) for c in cins], pixel_shuffle=ps, device=dev)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    f = 2 if ps else 1
    co = cout // 4 if ps else cout
    out = ops.View.alloc(Ho * f, Wo * f, co, dev)
    bufs = [[ops.View(torch.randn(H * W * c, device=dev), H, W, c, c) for c in cins] for _ in range(3)]
    extras = os.environ.get("CONV_BENCH_EXTRAS", "")      # any of: r (res1), s (res2), o (out2), l (input LeakyReLU): timing only
    kw = {}
    if extras:
        mk = lambda: ops.View(torch.randn(Ho * f * Wo * f * co, device=dev), Ho * f, Wo * f, co, co)
        if "r" in extras:
            kw["res1"] = mk()
        if "s" in extras:
            kw["res2"] = mk()
        if "o" in extras:
            kw["out2"], kw["slope2"] = mk(), 0.01
        if "l" in extras:
            kw["in_transform"], kw["in_slope"] = _lib.IN_LRELU, 0.01
    # ---- accuracy on the top rows (fp64 reference)
    ops.conv(pc, bufs[0], out, act=0.01, engine=engine)
    torch.cuda.synchronize()
    rows_in = min(H, check_rows * stride + k)
    x = torch.cat([v.as_tensor()[:rows_in].permute(2, 0, 1)[None] for v in bufs[0]], 1).double()
    ref = F.leaky_relu(F.conv2d(x, w.double().to(dev), b.double().to(dev), stride=stride, padding=pad), 0.01)
    if ps:
        ref = F.pixel_shuffle(ref, 2)
    rows_out = (min(check_rows, Ho) - k) * f          # rows not touched by the cut at the bottom of the crop
    got = out.as_tensor()[:rows_out].permute(2, 0, 1)[None].double()
    err = ((got - ref[:, :, :rows_out]).abs().max() / ref.abs().max()).item()
    if os.environ.get("CONV_BENCH_BIAS"):
        r = ref[:, :, :rows_out]
        d = got - r
        big = r.abs() > r.abs().mean()
        shrink = ((d * r.sign())[big] / r.abs()[big]).mean().item()      # < 0: results pulled towards zero
        rms = (d.pow(2).mean().sqrt() / r.pow(2).mean().sqrt()).item()
        steps = k * k * sum((c + 15) // 16 for c in cins)
        print(f"    [{engine}] {name}: signed relative error (towards zero < 0) {shrink:+.3e} = {shrink * 2 ** 24:+.2f} x 2^-24 "
              f"({shrink * 2 ** 24 / steps:+.3f} per K=16 accumulation step, {steps} steps), rms rel {rms:.3e}")
    # ---- time
    for i in range(3):
        ops.conv(pc, bufs[i], out, act=0.01, engine=engine, **kw)
    torch.cuda.synchronize()
    n = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops.conv(pc, bufs[i % 3], out, act=0.01, engine=engine, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * Ho * Wo * k * k * cin * cout
    return err, ms, flops / ms / 1e9


def main():
    only = os.environ.get("CONV_BENCH_ONLY")
    if only:
        SHAPES[:] = [s for s in SHAPES + EXTRA if only in s[0]]
    engines = sys.argv[1:] or ["h2", "simt"]
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    print(f"{'shape':26s} " + " ".join(f"{e + ' err':>11s} {e + ' ms':>9s} {e + ' TF/s':>9s}" for e in engines))
    for shape in SHAPES:
        cells = []
        for e in engines:
            try:
                err, ms, tf = run(e, shape, dev)
                cells.append(f"{err:11.2e} {ms:9.3f} {tf:9.1f}")
            except Exception as ex:  # noqa: BLE001
                cells.append(f"FAILED: {str(ex)[:120]}")
                if "CUDA" in str(ex) or "launch" in str(ex):
                    print(f"{shape[0]:26s} " + " ".join(cells), flush=True)
                    raise
        print(f"{shape[0]:26s} " + " ".join(cells), flush=True)


if __name__ == "__main__":
    main()
