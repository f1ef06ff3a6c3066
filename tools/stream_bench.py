"""--write_stream 1 at a named size (BASELINE.json config 4): I + N P-frames through encode_decode with real bitstream
files; prints per-frame wall times (the encoder pass + rANS encode, and the genuine decoder pass + rANS decode) and sizes.
usage: python tools/stream_bench.py [--size 1080p] [--p-frames 4] [--single-pass]"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--p-frames", type=int, default=4)
    ap.add_argument("--single-pass", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    from lssvc_b200 import IntraSS, LSSVC_extend, _lib
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    frames, (H, W) = bench.make_frames(bench.SIZES[args.size], 1 + args.p_frames, seed=0)
    net_i, net_p = IntraSS(seed=0).to(dev), LSSVC_extend(seed=1).to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
        n.single_pass_streams = args.single_pass
        n.update(force=True)
    tmp = tempfile.mkdtemp()
    dpb = None
    for idx, (b, e) in enumerate(frames):
        pb, pe = os.path.join(tmp, f"{idx}_bl.bin"), os.path.join(tmp, f"{idx}_el.bin")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if idx == 0:
            r = net_i.encode_decode(b.to(dev), e.to(dev), pb, pe, H // 2, W // 2, H, W)
            dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}
            extra = ""
        else:
            r = net_p.encode_decode(b.to(dev), e.to(dev), dpb, pb, pe, W, H, W // 2, H // 2)
            dpb = r["dpb"]
            extra = (f"  enc BL {r['encoding_time_BL'] * 1e3:6.1f} EL {r['encoding_time_EL'] * 1e3:6.1f}  "
                     f"dec BL {r['decoding_time_BL'] * 1e3:6.1f} EL {r['decoding_time_EL'] * 1e3:6.1f} ms")
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        dpb["ref_frame_bl"].clamp_(0, 1)
        dpb["ref_frame_el"].clamp_(0, 1)
        print(f"frame {idx} ({'I' if idx == 0 else 'P'}): {dt:8.1f} ms  bits BL {r['bit_bl']:9.0f} EL {r['bit_el']:9.0f}{extra}", flush=True)


if __name__ == "__main__":
    main()
