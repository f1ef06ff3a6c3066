"""HBM-bound kernels of the path (warps, resamplers, pools, element-wise, entropy) at their 1080p P-frame sizes: achieved
GB/s on ALGORITHMIC bytes (compulsory reads + writes at fp32, each tensor once) against the measured copy
bandwidth (MEASURED_PEAKS.json hbm_gbs).  CUDA events on the launching stream, inputs rotated over 3 buffer sets so
that nothing is re-read from L2.

usage: python tools/mem_bench.py            (prints a table; bench.py imports run() for its "hbm_kernels" object)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


WARM, REPS = 3, None       # tools/ncu_driver.py sets (1, 1): one warm-up and one measured launch per kernel


def _time(torch, fn, sets, n=12):
    n = REPS or n
    for i in range(WARM):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def run(torch, dev, H=1152, W=1920, peak_gbs=6544.7):
    from lssvc_b200 import ops
    V = ops.View
    rnd = lambda h, w, c, s=1.0: V(torch.randn(h * w * c, device=dev) * s, h, w, c, c)
    new = lambda h, w, c: V.alloc(h, w, c, dev)
    rows = []

    def add(name, sec, nbytes, note=""):
        gbs = nbytes / sec / 1e9
        rows.append({"kernel": name, "ms": round(sec * 1e3, 4), "bytes": int(nbytes), "gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 3),
                     "note": note})

    px = H * W
    # flow_warp: (2C*4 + 8) bytes per pixel
    for C, h, w in ((48, H, W), (64, H // 2, W // 2), (96, H // 4, W // 4)):
        sets = [(rnd(h, w, C), rnd(h, w, 2, 3.0), new(h, w, C)) for _ in range(3)]
        add(f"flow_warp C={C} {h}x{w}", _time(torch, lambda s: ops.flow_warp(s[0], s[1], s[2]), sets), (2 * C * 4 + 8) * h * w)
    # bilinear x2 up of a 64-channel feature (MvResampler / TextureResampler)
    sets = [(rnd(H // 2, W // 2, 64), new(H, W, 64)) for _ in range(3)]
    add("bilinear_resize x2 C=64 -> 1152x1920", _time(torch, lambda s: ops.bilinear_resize(s[0], s[1]), sets), (px // 4 + px) * 64 * 4)
    # pools
    sets = [(rnd(H, W, 32), new(H // 2, W // 2, 32)) for _ in range(3)]
    add("maxpool2 C=32 1152x1920", _time(torch, lambda s: ops.maxpool2(s[0], s[1]), sets), (px + px // 4) * 32 * 4)
    # softmax2_blend: 2-channel logits + a + b read, out written
    sets = [(rnd(H, W, 2), rnd(H, W, 48), rnd(H, W, 48), new(H, W, 48)) for _ in range(3)]
    add("softmax2_blend C=48 1152x1920", _time(torch, lambda s: ops.softmax2_blend(s[0], s[1], s[2], s[3]), sets), px * (2 + 48 * 3) * 4)
    # lrelu_copy
    sets = [(rnd(H // 2, W // 2, 64), new(H // 2, W // 2, 64)) for _ in range(3)]
    add("lrelu_copy C=64 576x960", _time(torch, lambda s: ops.lrelu_copy(s[0], 0.1, s[1]), sets), 2 * (px // 4) * 64 * 4)
    # depthwise 3x3 (unfused instances: 128 channels at 1/4)
    wdw, bdw = torch.randn(9, 128, device=dev), torch.randn(128, device=dev)
    sets = [(rnd(H // 4, W // 4, 128), new(H // 4, W // 4, 128)) for _ in range(3)]
    add("dwconv3x3 C=128 288x480", _time(torch, lambda s: ops.dwconv3x3(s[0], wdw, bdw, s[1]), sets), 2 * (px // 16) * 128 * 4)
    # OffsetDiversity fused (read 96ch@1/2 offsets + 48ch feature + 2ch flow, write 48ch)
    G, O, C = 16, 2, 48
    fw, fb = torch.randn(C, 2 * C // G, device=dev) * 0.1, torch.randn(C, device=dev)
    sets = [(rnd(H, W, C), rnd(H // 2, W // 2, 3 * G * O, 0.5), rnd(H, W, 2, 3.0), new(H, W, C)) for _ in range(2)]
    od_bytes = px * (C * 4 * 2 + 8) + (px // 4) * 96 * 4
    add("offset_diversity C=48 G=16 O=2 1152x1920 (i.i.d. offsets +-20 px, planar)",
        _time(torch, lambda s: ops.offset_diversity(s[0], s[1], s[2], fw, fb, G, O, 40.0, s[3], planar=True), sets, n=6), od_bytes,
        "worst case: every (pixel, group, offset) samples an unrelated position")
    add("offset_diversity (i.i.d. offsets, direct NHWC gather)",
        _time(torch, lambda s: ops.offset_diversity(s[0], s[1], s[2], fw, fb, G, O, 40.0, s[3], planar=False), sets, n=6), od_bytes)
    import torch.nn.functional as F
    smooth = []
    for s in sets:      # offsets as the conv stack produces them in a frame: smooth fields (low-pass noise), a few pixels of residue
        low = F.interpolate(torch.randn(1, 3 * G * O, H // 32, W // 32, device=dev), size=(H // 2, W // 2), mode="bicubic") * 0.1
        smooth.append((s[0], V(low[0].permute(1, 2, 0).contiguous().reshape(-1), H // 2, W // 2, 3 * G * O, 3 * G * O), s[2], s[3]))
    add("offset_diversity (smooth offsets, planar)",
        _time(torch, lambda s: ops.offset_diversity(s[0], s[1], s[2], fw, fb, G, O, 40.0, s[3], planar=True), smooth, n=6), od_bytes,
        "replaces a (32,3,H,W) grid_sample + 566 MB of grids + the grouped 1x1 conv")
    add("offset_diversity (smooth offsets, direct NHWC gather)",
        _time(torch, lambda s: ops.offset_diversity(s[0], s[1], s[2], fw, fb, G, O, 40.0, s[3], planar=False), smooth, n=6), od_bytes)
    # entropy: laplace quant + bits (y, mean, scale read; y_hat written) and the four-part step
    h, w = H // 16, W // 16
    bits = torch.zeros(2, dtype=torch.float64, device=dev)
    y, mean, sc, yh = rnd(h, w, 64, 4.0), rnd(h, w, 64), V(torch.rand(h * w * 64, device=dev) + 0.1, h, w, 64, 64), new(h, w, 64)
    add("laplace_quant 64x72x120 (+bits)", _time(torch, lambda s: ops.laplace_quant(y, mean, sc, None, yh, bits[1:2]), [0], n=50),
        h * w * 64 * 4 * 4, "launch-latency bound: 553 k symbols")
    y4, p8, yh4 = rnd(h, w, 128, 4.0), V(torch.rand(h * w * 256, device=dev) + 0.1, h, w, 256, 256), new(h, w, 128)
    add("four_part_step 128x72x120 (+bits)", _time(torch, lambda s: ops.four_part_step(y4, p8, 1, yh4, None, None, bits[1:2]), [0], n=50),
        h * w * (128 // 4) * 4 * 4, "launch-latency bound: one quarter (276 k symbols) per step")
    return rows


def main():
    import json
    import torch
    from lssvc_b200 import _lib
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    peak = 6544.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass
    rows = run(torch, dev, peak_gbs=peak)
    print(f"{'kernel':44s} {'ms':>8s} {'MB (alg.)':>10s} {'GB/s':>8s} {'of ' + str(peak):>10s}")
    for r in rows:
        print(f"{r['kernel']:44s} {r['ms']:8.4f} {r['bytes'] / 1e6:10.1f} {r['gbs']:8.1f} {100 * r['frac']:9.1f}%  {r['note']}")


if __name__ == "__main__":
    main()
