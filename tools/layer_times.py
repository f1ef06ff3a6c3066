"""Per-launch CUDA-event timing of one steady-state P-frame (no ncu): every ops.* call is bracketed by events on the
launching stream, then times are aggregated by conv shape, by module and by kernel family.

usage: python tools/layer_times.py [--size 1080p] [--engine h2]
Event timing serialises nothing (events are recorded in stream order), so the numbers are warm-cache and include
launch gaps only when the GPU runs ahead of the host.
"""
import argparse
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--engine", default=None)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    import torch
    import bench
    from lssvc_b200 import _lib, ops
    if args.engine:
        ops.set_engine(args.engine)
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    frames, shape_hr = bench.make_frames(bench.SIZES[args.size], 4, seed=0)
    coder = bench.Coder(dev, shape_hr)
    from lssvc_b200 import profile
    timer = None
    for idx, (b, e) in enumerate(frames):
        if idx == 3:
            timer = profile.LaunchTimer()
            timer.__enter__()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
        coder.step(idx, 100, b.to(dev), e.to(dev))
    t1.record()
    rows = timer.rows()
    timer.__exit__(None, None, None)
    total = t0.elapsed_time(t1)
    tsum = sum(r[2] for r in rows)
    print(f"P-frame wall {total:.2f} ms; sum of {len(rows)} bracketed launches {tsum:.2f} ms")
    fam = defaultdict(lambda: [0, 0.0])
    for n, i, t in rows:
        key = n if n != "conv" else f"conv[{i['engine']}]"
        fam[key][0] += 1
        fam[key][1] += t
    for k, (n, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:24s} {n:4d} {t:8.3f} ms {100 * t / tsum:5.1f}%")
    for fam_name in ("pw", "ffn"):
        g2 = defaultdict(lambda: [0, 0.0])
        for n, i, t in rows:
            if n == fam_name and i is not None:
                k = (i["cin"], i["cout"], i["Ho"], i["Wo"], i.get("dw", False), i.get("hidden", 0))
                g2[k][0] += 1
                g2[k][1] += t
        for (cin, cout, Ho, Wo, dw, hidden), (n, t) in sorted(g2.items(), key=lambda kv: -kv[1][1]):
            gb = 4.0 * Ho * Wo * (cin + cout) * n / 1e9
            print(f"  {fam_name:4s} {cin:4d}->{cout:4d} hidden {hidden:4d} dw={int(dw)} {Ho:5d}x{Wo:<5d} n={n:3d} {t:8.3f} ms  {gb / t * 1e3:6.0f} GB/s(min, in+out once)")
    convs = [(i, t) for n, i, t in rows if n == "conv"]
    groups = defaultdict(lambda: [0, 0.0, 0.0])
    for c, t in convs:
        key = (c["k"], c["stride"], c["cin"], c["cout"], c["Ho"], c["Wo"], c["ps"], c.get("extras", ""))
        g = groups[key]
        g[0] += 1
        g[1] += t
        g[2] += c["flops"]
    ctot = sum(t for _, t in convs)
    print(f"convs: {len(convs)} launches, {ctot:.2f} ms, {sum(c['flops'] for c, _ in convs) / ctot / 1e9:.1f} TFLOP/s")
    print(f"{'k':>2s} {'s':>1s} {'cin':>4s} {'cout':>4s} {'Ho':>5s} {'Wo':>5s} ps {'n':>3s} {'ms':>8s} {'share':>6s} {'TF/s':>7s} {'GB/s(min)':>9s}")
    for key, (n, t, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:args.top]:
        k, s, cin, cout, Ho, Wo, ps, ex = key
        gb = 4.0 * (Ho * s * Wo * s * cin + Ho * Wo * cout * (1 + sum(ex.count(x) for x in "rso"))) * n / 1e9
        print(f"{k:2d} {s:1d} {cin:4d} {cout:4d} {Ho:5d} {Wo:5d} {int(ps):2d} {n:3d} {t:8.3f} {100 * t / ctot:5.1f}% {fl / t / 1e9:7.1f} {gb / t * 1e3:9.0f} {ex}")
    mods = defaultdict(lambda: [0, 0.0, 0.0])
    for c, t in convs:
        m = re.sub(r"^base_layer_model\.", "BL.", c["name"] or "?")
        m = ".".join(m.split(".")[:2 if m.startswith("BL.") else 1])
        mods[m][0] += 1
        mods[m][1] += t
        mods[m][2] += c["flops"]
    print()
    for m, (n, t, fl) in sorted(mods.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"{m:40s} {n:4d} {t:8.3f} ms {100 * t / ctot:5.1f}% {fl / t / 1e9:7.1f} TF/s")


if __name__ == "__main__":
    main()
