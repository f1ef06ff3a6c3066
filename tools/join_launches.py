"""Join an ncu launch list of `tools/profile_frame.py` with the host-side conv trace (walked here on CPU in dry-run
mode at the same size), so every tensor-core conv launch gets its layer name, shape and achieved TFLOP/s.
usage: python tools/join_launches.py launches.csv [size] [p_frames]"""
import csv
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def trace(size, p_frames):
    import torch
    import bench
    from lssvc_b200 import IntraSS, LSSVC_extend, _lib, ops, synth
    _lib.DRY_RUN = True
    pad = synth.interlayer_padding(*bench.SIZES[size], 2.0)
    H, W = pad["HR_padded_size"]
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    x_bl, x_el = torch.zeros(1, 3, H // 2, W // 2), torch.zeros(1, 3, H, W)
    ops.TRACE = []
    frames = []
    r = net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
    frames.append(ops.TRACE)
    dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}
    for _ in range(p_frames):
        ops.TRACE = []
        r = net_p.encode_decode(x_bl, x_el, dpb, None, None, W, H, W // 2, H // 2)
        dpb = r["dpb"]
        frames.append(ops.TRACE)
    return frames


def main():
    path = sys.argv[1]
    size = sys.argv[2] if len(sys.argv) > 2 else "1080p"
    p_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    frames = trace(size, p_frames)
    lines = [l for l in open(path) if l.startswith('"')]
    tc = [float(r["Metric Value"].replace(",", "")) for r in csv.DictReader(lines)
          if r.get("Metric Name") == "gpu__time_duration.sum" and ("conv_tc" in r["Kernel Name"] or "conv_h2" in r["Kernel Name"])]
    convs = [c for f in frames for c in f if c["engine"] != "simt"]
    assert len(convs) == len(tc), (len(convs), len(tc))
    n_last = len([c for c in frames[-1] if c["engine"] != "simt"])
    last = list(zip(convs[-n_last:], tc[-n_last:]))
    tot_ns = sum(t for _, t in last)
    tot_fl = sum(c["flops"] for c, _ in last)
    print(f"last P-frame: {n_last} tensor-core convs, {tot_ns / 1e6:.2f} ms, {tot_fl / 1e12:.3f} TFLOP -> {tot_fl / tot_ns / 1e3:.1f} TFLOP/s")
    groups = defaultdict(lambda: [0, 0.0, 0.0])
    for c, t in last:
        key = (c["k"], c["stride"], c["cin"], c["cout"], c["Ho"], c["Wo"])
        g = groups[key]
        g[0] += 1
        g[1] += t
        g[2] += c["flops"]
    print(f"{'k':>2s} {'s':>1s} {'cin':>4s} {'cout':>4s} {'Ho':>5s} {'Wo':>5s} {'n':>3s} {'ms':>8s} {'share':>6s} {'TF/s':>7s}")
    for key, (n, ns, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{key[0]:2d} {key[1]:1d} {key[2]:4d} {key[3]:4d} {key[4]:5d} {key[5]:5d} {n:3d} {ns / 1e6:8.3f} {100 * ns / tot_ns:5.1f}% {fl / ns / 1e3:7.1f}")
    mods = defaultdict(lambda: [0, 0.0, 0.0])
    for c, t in last:
        m = re.sub(r"^base_layer_model\.", "BL.", c["name"] or "?")
        m = ".".join(m.split(".")[:2 if m.startswith("BL.") else 1])
        mods[m][0] += 1
        mods[m][1] += t
        mods[m][2] += c["flops"]
    print()
    for m, (n, ns, fl) in sorted(mods.items(), key=lambda kv: -kv[1][1]):
        print(f"{m:40s} {n:4d} {ns / 1e6:8.3f} ms {100 * ns / tot_ns:5.1f}% {fl / ns / 1e3:7.1f} TF/s")


if __name__ == "__main__":
    main()
