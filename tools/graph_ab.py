"""A/B of whole-frame CUDA graphs: ms per steady-state P-frame (host wall clock, each frame ends with its own bit-count
sync) with LSSVC.use_graphs off and on, same frames.   usage: python tools/graph_ab.py [--size 1080p] [--frames 10]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--frames", type=int, default=10)
    args = ap.parse_args()
    import torch
    import bench
    from lssvc_b200 import _lib
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().lssvc_device_check(0), "device_check")
    n = args.frames + 4
    frames, shape_hr = bench.make_frames(bench.SIZES[args.size], n, seed=0)
    devf = [(b.to(dev), e.to(dev)) for b, e in frames]
    for mode in (False, True, False, True):
        coder = bench.Coder(dev, shape_hr)
        coder.net_p.use_graphs = mode
        times = []
        for i in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            coder.step(i, 1000, *devf[i])
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        tail = sorted(times[4:])
        print(f"graphs={mode!s:5s}  P-frame ms: median {tail[len(tail) // 2]:.2f}  min {tail[0]:.2f}  max {tail[-1]:.2f}   first frames "
              + " ".join(f"{t:.0f}" for t in times[:4]), flush=True)
        del coder
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
