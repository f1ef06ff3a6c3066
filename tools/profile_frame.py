"""Runs I + N P-frames of the two-layer coder at a named size on cuda:0 (no oracle, no timing): the command the ncu
launch lists under profiles/ are taken from.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_frame.py --size 1080p --p-frames 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--p-frames", type=int, default=2)
    ap.add_argument("--engine", default=None)
    args = ap.parse_args()
    import torch
    import bench
    from lssvc_b200 import _lib, ops
    if args.engine:
        ops.set_engine(args.engine)
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    frames, shape_hr = bench.make_frames(bench.SIZES[args.size], 1 + args.p_frames, seed=0)
    coder = bench.Coder(dev, shape_hr)
    for idx, (b, e) in enumerate(frames):
        l0 = _lib.launch_count()
        bb, be, _, _ = coder.step(idx, 1 + args.p_frames, b.to(dev), e.to(dev))
        torch.cuda.synchronize()
        print(f"frame {idx}: {_lib.launch_count() - l0} launches, bits {bb:.0f}/{be:.0f}", flush=True)


if __name__ == "__main__":
    main()
